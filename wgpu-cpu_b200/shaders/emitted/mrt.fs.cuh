// wgsl2cuda: stage=fragment entry=fs_main
namespace wgb_fragment {
struct Camera { mat4x4f matrix; };
struct Interstage { vec4f clip; vec4f tint; };
struct Targets { vec4f first; vec4f third; vec4f second; };
struct WgbInvocation {
    bool killed = false;
};
WGB_DEV Targets fs_main(const WgbDraw& wgb, WgbInvocation& wgb_inv, Interstage frag);
WGB_DEV Targets fs_main(const WgbDraw& wgb, WgbInvocation& wgb_inv, Interstage frag) {
    const vec4f third = vec4f(frag.clip.z, wgb_mul(frag.tint.x, frag.tint.y), 0.25f, 1.0f);
    const vec4f second = vec4f(frag.tint.z, frag.tint.y, frag.tint.x, 0.5f);
    return Targets{frag.tint, third, second};
}
}  // namespace wgb_fragment
#define WGB_FS_COLOR_MASK 7
#define WGB_FS_WRITES_FRAG_DEPTH 0
#define WGB_FS_MAY_DISCARD 0
#define WGB_FS_EARLY_DEPTH 0
#define WGB_FS_USES_FRONT_FACING 0
WGB_DEV constexpr int wgb_fs_interp(int slot) {
    return (slot >= WGB_VS_LOC0_SLOT && slot < WGB_VS_LOC0_SLOT + 4) ? 1 : 0;
}
WGB_DEV bool wgb_fs_entry(const WgbDraw& wgb, const WgbFragIn& fi, const u32* vary, WgbFragOut& out) {
    wgb_fragment::WgbInvocation wgb_inv;
    wgb_fragment::Interstage a0;
    a0.clip = fi.position;
    a0.tint = wgb_get<vec4f>(vary, WGB_VS_LOC0_SLOT);
    const wgb_fragment::Targets r = wgb_fragment::fs_main(wgb, wgb_inv, a0);
    if (wgb_inv.killed) return false;
    out.color[0] = r.first;
    out.color[2] = r.third;
    out.color[1] = r.second;
    return true;
}
