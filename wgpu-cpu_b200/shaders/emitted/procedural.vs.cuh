// wgsl2cuda: stage=vertex entry=vs_main
namespace wgb_vertex {
struct VertexInput { u32 vertex_index; };
struct VertexOutput { vec4f position; };
struct WgbInvocation {
    bool killed = false;
};
WGB_DEV VertexOutput vs_main(const WgbDraw& wgb, WgbInvocation& wgb_inv, VertexInput input);
WGB_DEV VertexOutput vs_main(const WgbDraw& wgb, WgbInvocation& wgb_inv, VertexInput input) {
    const u32 vi = wgb_irem(input.vertex_index, 6u);
    const f32 ux = wgb_to_f32((((vi == 1u) || (vi == 4u)) || (vi == 5u)));
    const f32 uy = wgb_to_f32((((vi == 2u) || (vi == 3u)) || (vi == 5u)));
    const vec4f position = vec4f(wgb_sub(wgb_mul(ux, 2.0f), 1.0f), wgb_sub(wgb_mul(uy, 2.0f), 1.0f), 0.0f, 1.0f);
    return VertexOutput{position};
}
}  // namespace wgb_vertex
#define WGB_VS_VARYING_SLOTS 0
WGB_DEV void wgb_vs_entry(const WgbDraw& wgb, u32 vertex_index, u32 instance_index, vec4f& position, u32* vary, u32& oob) {
    wgb_vertex::WgbInvocation wgb_inv;
    wgb_vertex::VertexInput a0;
    a0.vertex_index = vertex_index;
    const wgb_vertex::VertexOutput r = wgb_vertex::vs_main(wgb, wgb_inv, a0);
    position = r.position;
}
