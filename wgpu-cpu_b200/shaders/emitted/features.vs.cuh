// wgsl2cuda: stage=vertex entry=vs_main
namespace wgb_vertex {
struct VertexInput { u32 vertex_index; u32 instance_index; vec4f vertex_position; vec4f vertex_color; };
struct VertexOutput { vec4f position; vec4f color; u32 tag; };
struct Params { mat4x4f matrix; vec4f instance_offset; };
struct FragmentInput { vec4f position; bool front_facing; vec4f color; u32 tag; };
struct WgbInvocation {
    bool killed = false;
};
WGB_DEV VertexOutput vs_main(const WgbDraw& wgb, WgbInvocation& wgb_inv, VertexInput input);
WGB_DEV VertexOutput vs_main(const WgbDraw& wgb, WgbInvocation& wgb_inv, VertexInput input) {
    const f32 shift = wgb_to_f32(input.instance_index);
    const vec4f p = (wgb_load<mat4x4f>(wgb, 0, 0, 0u) * input.vertex_position);
    const vec4f position = vec4f(wgb_add(p.x, wgb_mul(wgb_load<f32>(wgb, 0, 0, 64u), shift)), wgb_add(p.y, wgb_mul(wgb_load<f32>(wgb, 0, 0, 68u), shift)), wgb_add(p.z, wgb_mul(wgb_load<f32>(wgb, 0, 0, 72u), shift)), p.w);
    return VertexOutput{position, input.vertex_color, (input.vertex_index + (input.instance_index * 1000u))};
}
}  // namespace wgb_vertex
#define WGB_VS_VARYING_SLOTS 5
#define WGB_VS_LOC0_SLOT 0
#define WGB_VS_LOC1_SLOT 4
WGB_DEV void wgb_vs_entry(const WgbDraw& wgb, u32 vertex_index, u32 instance_index, vec4f& position, u32* vary, u32& oob) {
    wgb_vertex::WgbInvocation wgb_inv;
    wgb_vertex::VertexInput a0;
    a0.vertex_index = vertex_index;
    a0.instance_index = instance_index;
    a0.vertex_position = WGB_FETCH(vec4f, 0);
    a0.vertex_color = WGB_FETCH(vec4f, 1);
    const wgb_vertex::VertexOutput r = wgb_vertex::vs_main(wgb, wgb_inv, a0);
    position = r.position;
    wgb_put(vary, WGB_VS_LOC0_SLOT, r.color);
    wgb_put(vary, WGB_VS_LOC1_SLOT, r.tag);
}
