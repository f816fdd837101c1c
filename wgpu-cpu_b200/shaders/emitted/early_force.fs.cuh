// wgsl2cuda: stage=fragment entry=fs_main
namespace wgb_fragment {
struct Camera { mat4x4f matrix; };
struct Interstage { vec4f clip; vec4f tint; };
struct Shaded { f32 depth; vec4f tint; };
struct WgbInvocation {
    bool killed = false;
};
WGB_DEV Shaded fs_main(const WgbDraw& wgb, WgbInvocation& wgb_inv, Interstage frag);
WGB_DEV Shaded fs_main(const WgbDraw& wgb, WgbInvocation& wgb_inv, Interstage frag) {
    if ((frag.tint.y > 0.600000024f)) {
        { wgb_inv.killed = true; return Shaded(); }
    }
    return Shaded{wgb_mul(frag.clip.z, 0.5f), frag.tint};
}
}  // namespace wgb_fragment
#define WGB_FS_COLOR_MASK 1
#define WGB_FS_WRITES_FRAG_DEPTH 1
#define WGB_FS_MAY_DISCARD 1
#define WGB_FS_EARLY_DEPTH 1
#define WGB_FS_USES_FRONT_FACING 0
WGB_DEV constexpr int wgb_fs_interp(int slot) {
    return (slot >= WGB_VS_LOC0_SLOT && slot < WGB_VS_LOC0_SLOT + 4) ? 1 : 0;
}
WGB_DEV bool wgb_fs_entry(const WgbDraw& wgb, const WgbFragIn& fi, const u32* vary, WgbFragOut& out) {
    wgb_fragment::WgbInvocation wgb_inv;
    wgb_fragment::Interstage a0;
    a0.clip = fi.position;
    a0.tint = wgb_get<vec4f>(vary, WGB_VS_LOC0_SLOT);
    const wgb_fragment::Shaded r = wgb_fragment::fs_main(wgb, wgb_inv, a0);
    if (wgb_inv.killed) return false;
    out.frag_depth = r.depth;
    out.color[0] = r.tint;
    return true;
}
