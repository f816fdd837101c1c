// wgsl2cuda: stage=fragment entry=fs_main
namespace wgb_fragment {
struct VertexInput { u32 vertex_index; };
struct VertexOutput { vec4f position; };
struct WgbInvocation {
    bool killed = false;
};
WGB_DEV vec4f fs_main(const WgbDraw& wgb, WgbInvocation& wgb_inv, VertexOutput input);
WGB_DEV vec4f fs_main(const WgbDraw& wgb, WgbInvocation& wgb_inv, VertexOutput input) {
    const f32 cx = wgb_sub(wgb_div(input.position.x, 2560.0f), 2.0f);
    const f32 cy = wgb_sub(wgb_div(input.position.y, 2160.0f), 1.0f);
    f32 zx = 0.0f;
    f32 zy = 0.0f;
    f32 acc = 0.0f;
    {
        i32 i = 0;
        for (; (i < 64); i = (i + 1)) {
            const f32 nx = wgb_add(wgb_sub(wgb_mul(zx, zx), wgb_mul(zy, zy)), cx);
            const f32 ny = wgb_add(wgb_mul(wgb_mul(2.0f, zx), zy), cy);
            const f32 r2 = wgb_add(wgb_mul(nx, nx), wgb_mul(ny, ny));
            const bool escaped = (r2 > 4.0f);
            zx = wgb_select(nx, zx, escaped);
            zy = wgb_select(ny, zy, escaped);
            acc = wgb_add(acc, wgb_select(1.0f, 0.0f, escaped));
        }
    }
    const f32 t = wgb_div(acc, 64.0f);
    return vec4f(t, wgb_mul(t, t), wgb_sub(1.0f, t), 1.0f);
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
    wgb_fragment::VertexOutput a0;
    a0.position = fi.position;
    const vec4f r = wgb_fragment::fs_main(wgb, wgb_inv, a0);
    if (wgb_inv.killed) return false;
    out.color[0] = r;
    return true;
}
