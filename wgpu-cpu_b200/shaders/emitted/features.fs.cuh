// wgsl2cuda: stage=fragment entry=fs_main
namespace wgb_fragment {
struct VertexInput { u32 vertex_index; u32 instance_index; vec4f vertex_position; vec4f vertex_color; };
struct VertexOutput { vec4f position; vec4f color; u32 tag; };
struct Params { mat4x4f matrix; vec4f instance_offset; };
struct FragmentInput { vec4f position; bool front_facing; vec4f color; u32 tag; };
struct WgbInvocation {
    bool killed = false;
};
WGB_DEV vec4f fs_main(const WgbDraw& wgb, WgbInvocation& wgb_inv, FragmentInput input);
WGB_DEV vec4f fs_main(const WgbDraw& wgb, WgbInvocation& wgb_inv, FragmentInput input) {
    const f32 lx = wgb_select(input.position.x, 0.0f, (input.position.x < 0.0f));
    const f32 ly = wgb_select(input.position.y, 0.0f, (input.position.y < 0.0f));
    const u32 px = wgb_to_u32(wgb_select(lx, 4096.0f, (lx > 4096.0f)));
    const u32 py = wgb_to_u32(wgb_select(ly, 4096.0f, (ly > 4096.0f)));
    if ((wgb_irem((wgb_idiv(px, 4u) + wgb_idiv(py, 4u)), 3u) == 0u)) {
        { wgb_inv.killed = true; return vec4f(); }
    }
    const f32 t = wgb_div(wgb_to_f32(wgb_irem(input.tag, 7u)), 7.0f);
    const f32 facing = wgb_to_f32(input.front_facing);
    return vec4f(input.color.x, wgb_mul(input.color.y, t), wgb_mul(input.color.z, facing), 1.0f);
}
}  // namespace wgb_fragment
#define WGB_FS_COLOR_MASK 1
#define WGB_FS_WRITES_FRAG_DEPTH 0
#define WGB_FS_MAY_DISCARD 1
#define WGB_FS_EARLY_DEPTH 0
#define WGB_FS_USES_FRONT_FACING 1
WGB_DEV constexpr int wgb_fs_interp(int slot) {
    return (slot >= WGB_VS_LOC0_SLOT && slot < WGB_VS_LOC0_SLOT + 4) ? 1 : (slot >= WGB_VS_LOC1_SLOT && slot < WGB_VS_LOC1_SLOT + 1) ? 0 : 0;
}
WGB_DEV bool wgb_fs_entry(const WgbDraw& wgb, const WgbFragIn& fi, const u32* vary, WgbFragOut& out) {
    wgb_fragment::WgbInvocation wgb_inv;
    wgb_fragment::FragmentInput a0;
    a0.position = fi.position;
    a0.front_facing = fi.front_facing;
    a0.color = wgb_get<vec4f>(vary, WGB_VS_LOC0_SLOT);
    a0.tag = wgb_get<u32>(vary, WGB_VS_LOC1_SLOT);
    const vec4f r = wgb_fragment::fs_main(wgb, wgb_inv, a0);
    if (wgb_inv.killed) return false;
    out.color[0] = r;
    return true;
}
