// wgsl2cuda: stage=vertex entry=vs_main
namespace wgb_vertex {
struct VertexInput { u32 vertex_index; };
struct VertexOutput { vec4f position; vec4f color; };
struct WgbInvocation {
    bool killed = false;
};
WGB_DEV VertexOutput vs_main(const WgbDraw& wgb, WgbInvocation& wgb_inv, VertexInput input);
WGB_DEV VertexOutput vs_main(const WgbDraw& wgb, WgbInvocation& wgb_inv, VertexInput input) {
    const u32 vertex_index = wgb_irem(input.vertex_index, 3u);
    const f32 x = wgb_to_f32((wgb_to_i32(vertex_index) - 1));
    const f32 y = wgb_to_f32(((wgb_to_i32((vertex_index & 1u)) * 2) - 1));
    const vec4f position = vec4f(x, y, 0.0f, 1.0f);
    const f32 r = wgb_to_f32((vertex_index == 0u));
    const f32 g = wgb_to_f32((vertex_index == 1u));
    const f32 b = wgb_to_f32((vertex_index == 2u));
    const vec4f color = vec4f(r, g, b, 1.0f);
    return VertexOutput{position, color};
}
}  // namespace wgb_vertex
#define WGB_VS_VARYING_SLOTS 4
#define WGB_VS_LOC0_SLOT 0
WGB_DEV void wgb_vs_entry(const WgbDraw& wgb, u32 vertex_index, u32 instance_index, vec4f& position, u32* vary, u32& oob) {
    wgb_vertex::WgbInvocation wgb_inv;
    wgb_vertex::VertexInput a0;
    a0.vertex_index = vertex_index;
    const wgb_vertex::VertexOutput r = wgb_vertex::vs_main(wgb, wgb_inv, a0);
    position = r.position;
    wgb_put(vary, WGB_VS_LOC0_SLOT, r.color);
}
