// wgsl2cuda: stage=fragment entry=fs_main
namespace wgb_fragment {
struct Camera { mat4x4f matrix; };
struct WgbInvocation {
    bool killed = false;
};
WGB_DEV void fs_main(const WgbDraw& wgb, WgbInvocation& wgb_inv);
WGB_DEV void fs_main(const WgbDraw& wgb, WgbInvocation& wgb_inv) {
}
}  // namespace wgb_fragment
#define WGB_FS_COLOR_MASK 0
#define WGB_FS_WRITES_FRAG_DEPTH 0
#define WGB_FS_MAY_DISCARD 0
#define WGB_FS_EARLY_DEPTH 1
#define WGB_FS_USES_FRONT_FACING 0
WGB_DEV constexpr int wgb_fs_interp(int slot) {
    return 0;
}
WGB_DEV bool wgb_fs_entry(const WgbDraw& wgb, const WgbFragIn& fi, const u32* vary, WgbFragOut& out) {
    wgb_fragment::WgbInvocation wgb_inv;
    wgb_fragment::fs_main(wgb, wgb_inv);
    if (wgb_inv.killed) return false;
    return true;
}
