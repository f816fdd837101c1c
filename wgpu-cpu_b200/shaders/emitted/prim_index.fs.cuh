// wgsl2cuda: stage=fragment entry=fs_main
namespace wgb_fragment {
struct Camera { mat4x4f matrix; };
struct WgbInvocation {
    bool killed = false;
};
WGB_DEV vec4f fs_main(const WgbDraw& wgb, WgbInvocation& wgb_inv, u32 primitive, u32 sample, u32 mask);
WGB_DEV vec4f fs_main(const WgbDraw& wgb, WgbInvocation& wgb_inv, u32 primitive, u32 sample, u32 mask) {
    const f32 low = wgb_div(wgb_to_f32((primitive & 255u)), 255.0f);
    const f32 high = wgb_div(wgb_to_f32(((primitive >> ((u32)(8u) & 31u)) & 255u)), 255.0f);
    return vec4f(low, high, wgb_add(wgb_to_f32(sample), wgb_mul(wgb_to_f32((mask & 1u)), 0.5f)), 1.0f);
}
}  // namespace wgb_fragment
#define WGB_FS_COLOR_MASK 1
#define WGB_FS_WRITES_FRAG_DEPTH 0
#define WGB_FS_MAY_DISCARD 0
#define WGB_FS_EARLY_DEPTH 0
#define WGB_FS_USES_FRONT_FACING 0
WGB_DEV constexpr int wgb_fs_interp(int slot) {
    return 0;
}
WGB_DEV bool wgb_fs_entry(const WgbDraw& wgb, const WgbFragIn& fi, const u32* vary, WgbFragOut& out) {
    wgb_fragment::WgbInvocation wgb_inv;
    u32 a0;
    a0 = fi.primitive_index;
    u32 a1;
    a1 = fi.sample_index;
    u32 a2;
    a2 = fi.sample_mask;
    const vec4f r = wgb_fragment::fs_main(wgb, wgb_inv, a0, a1, a2);
    if (wgb_inv.killed) return false;
    out.color[0] = r;
    return true;
}
