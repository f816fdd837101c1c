// wgsl2cuda: stage=vertex entry=vs_main
namespace wgb_vertex {
struct Corner { vec4f clip; vec4f tint; };
struct WgbInvocation {
    bool killed = false;
};
WGB_DEV Corner vs_main(const WgbDraw& wgb, WgbInvocation& wgb_inv, u32 index);
WGB_DEV Corner vs_main(const WgbDraw& wgb, WgbInvocation& wgb_inv, u32 index) {
    const u32 k = wgb_irem(index, 3u);
    const vec4f clip = vec4f(wgb_to_f32((wgb_to_i32(k) - 1)), wgb_to_f32(((wgb_to_i32((k & 1u)) * 2) - 1)), 0.0f, 1.0f);
    const vec4f tint = vec4f(wgb_to_f32((k == 0u)), wgb_to_f32((k == 1u)), wgb_to_f32((k == 2u)), 1.0f);
    return Corner{clip, tint};
}
}  // namespace wgb_vertex
#define WGB_VS_VARYING_SLOTS 4
#define WGB_VS_LOC0_SLOT 0
WGB_DEV void wgb_vs_entry(const WgbDraw& wgb, u32 vertex_index, u32 instance_index, vec4f& position, u32* vary, u32& oob) {
    wgb_vertex::WgbInvocation wgb_inv;
    u32 a0;
    a0 = vertex_index;
    const wgb_vertex::Corner r = wgb_vertex::vs_main(wgb, wgb_inv, a0);
    position = r.clip;
    wgb_put(vary, WGB_VS_LOC0_SLOT, r.tint);
}
