// wgsl2cuda: stage=vertex entry=vs_main
namespace wgb_vertex {
struct VertexInput { u32 vertex_index; vec4f vertex_position; vec2f uv; };
struct VertexOutput { vec4f position; vec2f uv; };
struct Camera { mat4x4f matrix; };
struct WgbInvocation {
    bool killed = false;
};
WGB_DEV VertexOutput vs_main(const WgbDraw& wgb, WgbInvocation& wgb_inv, VertexInput input);
WGB_DEV VertexOutput vs_main(const WgbDraw& wgb, WgbInvocation& wgb_inv, VertexInput input) {
    const vec4f position = (wgb_load<mat4x4f>(wgb, 0, 0, 0u) * input.vertex_position);
    return VertexOutput{position, input.uv};
}
}  // namespace wgb_vertex
#define WGB_VS_VARYING_SLOTS 2
#define WGB_VS_LOC0_SLOT 0
WGB_DEV void wgb_vs_entry(const WgbDraw& wgb, u32 vertex_index, u32 instance_index, vec4f& position, u32* vary, u32& oob) {
    wgb_vertex::WgbInvocation wgb_inv;
    wgb_vertex::VertexInput a0;
    a0.vertex_index = vertex_index;
    a0.vertex_position = WGB_FETCH(vec4f, 0);
    a0.uv = WGB_FETCH(vec2f, 1);
    const wgb_vertex::VertexOutput r = wgb_vertex::vs_main(wgb, wgb_inv, a0);
    position = r.position;
    wgb_put(vary, WGB_VS_LOC0_SLOT, r.uv);
}
