// wgsl2cuda: stage=fragment entry=fs_main
namespace wgb_fragment {
struct Camera { mat4x4f matrix; };
struct Interstage { vec4f clip; vec4f corrected; vec2f screen_space; f32 provoking; };
struct WgbInvocation {
    bool killed = false;
};
WGB_DEV vec4f fs_main(const WgbDraw& wgb, WgbInvocation& wgb_inv, Interstage frag);
WGB_DEV vec4f fs_main(const WgbDraw& wgb, WgbInvocation& wgb_inv, Interstage frag) {
    return vec4f(frag.corrected.x, wgb_add(wgb_mul(frag.corrected.y, 0.5f), wgb_mul(frag.screen_space.y, 0.5f)), frag.screen_space.x, frag.provoking);
}
}  // namespace wgb_fragment
#define WGB_FS_COLOR_MASK 1
#define WGB_FS_WRITES_FRAG_DEPTH 0
#define WGB_FS_MAY_DISCARD 0
#define WGB_FS_EARLY_DEPTH 0
#define WGB_FS_USES_FRONT_FACING 0
WGB_DEV constexpr int wgb_fs_interp(int slot) {
    return (slot >= WGB_VS_LOC0_SLOT && slot < WGB_VS_LOC0_SLOT + 4) ? 2 : (slot >= WGB_VS_LOC1_SLOT && slot < WGB_VS_LOC1_SLOT + 2) ? 1 : (slot >= WGB_VS_LOC2_SLOT && slot < WGB_VS_LOC2_SLOT + 1) ? 0 : 0;
}
WGB_DEV bool wgb_fs_entry(const WgbDraw& wgb, const WgbFragIn& fi, const u32* vary, WgbFragOut& out) {
    wgb_fragment::WgbInvocation wgb_inv;
    wgb_fragment::Interstage a0;
    a0.clip = fi.position;
    a0.corrected = wgb_get<vec4f>(vary, WGB_VS_LOC0_SLOT);
    a0.screen_space = wgb_get<vec2f>(vary, WGB_VS_LOC1_SLOT);
    a0.provoking = wgb_get<f32>(vary, WGB_VS_LOC2_SLOT);
    const vec4f r = wgb_fragment::fs_main(wgb, wgb_inv, a0);
    if (wgb_inv.killed) return false;
    out.color[0] = r;
    return true;
}
