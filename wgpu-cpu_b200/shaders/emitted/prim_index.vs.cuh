// wgsl2cuda: stage=vertex entry=vs_main
namespace wgb_vertex {
struct Camera { mat4x4f matrix; };
struct WgbInvocation {
    bool killed = false;
};
WGB_DEV vec4f vs_main(const WgbDraw& wgb, WgbInvocation& wgb_inv, vec4f object_position, vec4f tint);
WGB_DEV vec4f vs_main(const WgbDraw& wgb, WgbInvocation& wgb_inv, vec4f object_position, vec4f tint) {
    return (wgb_load<mat4x4f>(wgb, 0, 0, 0u) * object_position);
}
}  // namespace wgb_vertex
#define WGB_VS_VARYING_SLOTS 0
WGB_DEV void wgb_vs_entry(const WgbDraw& wgb, u32 vertex_index, u32 instance_index, vec4f& position, u32* vary, u32& oob) {
    wgb_vertex::WgbInvocation wgb_inv;
    vec4f a0;
    a0 = WGB_FETCH(vec4f, 0);
    vec4f a1;
    a1 = WGB_FETCH(vec4f, 1);
    const vec4f r = wgb_vertex::vs_main(wgb, wgb_inv, a0, a1);
    position = r;
}
