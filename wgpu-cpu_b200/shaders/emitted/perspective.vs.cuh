// wgsl2cuda: stage=vertex entry=vs_main
namespace wgb_vertex {
struct Camera { mat4x4f matrix; };
struct Interstage { vec4f clip; vec4f corrected; vec2f screen_space; f32 provoking; };
struct WgbInvocation {
    bool killed = false;
};
WGB_DEV Interstage vs_main(const WgbDraw& wgb, WgbInvocation& wgb_inv, vec4f object_position, vec4f tint);
WGB_DEV Interstage vs_main(const WgbDraw& wgb, WgbInvocation& wgb_inv, vec4f object_position, vec4f tint) {
    return Interstage{(wgb_load<mat4x4f>(wgb, 0, 0, 0u) * object_position), tint, vec2f(tint.x, tint.y), tint.z};
}
}  // namespace wgb_vertex
#define WGB_VS_VARYING_SLOTS 7
#define WGB_VS_LOC0_SLOT 0
#define WGB_VS_LOC1_SLOT 4
#define WGB_VS_LOC2_SLOT 6
WGB_DEV void wgb_vs_entry(const WgbDraw& wgb, u32 vertex_index, u32 instance_index, vec4f& position, u32* vary, u32& oob) {
    wgb_vertex::WgbInvocation wgb_inv;
    vec4f a0;
    a0 = WGB_FETCH(vec4f, 0);
    vec4f a1;
    a1 = WGB_FETCH(vec4f, 1);
    const wgb_vertex::Interstage r = wgb_vertex::vs_main(wgb, wgb_inv, a0, a1);
    position = r.clip;
    wgb_put(vary, WGB_VS_LOC0_SLOT, r.corrected);
    wgb_put(vary, WGB_VS_LOC1_SLOT, r.screen_space);
    wgb_put(vary, WGB_VS_LOC2_SLOT, r.provoking);
}
