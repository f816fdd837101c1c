/*
 * oracle_shaders.cpp -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Hand-written C++ restatements of the WGSL shaders used by the BASELINE configs and
 * the parity tests, in the operation order wgpu-cpu's shader JIT (naga-cranelift)
 * emits: one IEEE f32 operation per WGSL operator, no FMA contraction
 * (naga-cranelift/src/expression/binary.rs:420-449), mat4x4f * vec4f as
 * ((v0*c0 + v1*c1) + v2*c2) + v3*c3 over columns (binary.rs:297-323), casts per
 * expression/as.rs:60-184.  These are written independently of the product's
 * WGSL->CUDA emitter so that the two can be checked against each other.
 */
#include "oracle_internal.h"

namespace orc {

static inline Vec4 load_vec4(const uint8_t* p) { Vec4 v; std::memcpy(&v, p, 16); return v; }
static inline void store_vec4(uint8_t* p, Vec4 v) { std::memcpy(p, &v, 16); }
static inline void store_f32(uint8_t* p, float v) { std::memcpy(p, &v, 4); }
static inline float load_f32(const uint8_t* p) { float v; std::memcpy(&v, p, 4); return v; }
static inline uint32_t load_u32(const uint8_t* p) { uint32_t v; std::memcpy(&v, p, 4); return v; }
static inline void store_u32(uint8_t* p, uint32_t v) { std::memcpy(p, &v, 4); }

/* binary.rs:297-323: column_sum = v0*col0; column_sum += v_i*col_i (separate fmul, fadd) */
static inline Vec4 mat4_mul_vec4(const float m[16], Vec4 v) {
    const float vv[4] = {v.x, v.y, v.z, v.w};
    float s[4];
    for (int r = 0; r < 4; r++) s[r] = vv[0] * m[r];
    for (int i = 1; i < 4; i++)
        for (int r = 0; r < 4; r++) {
            const float x = vv[i] * m[4 * i + r];
            s[r] = s[r] + x;
        }
    return {s[0], s[1], s[2], s[3]};
}

static inline const uint8_t* need_attr(const VsIn& in, uint32_t loc, uint32_t size, int* err) {
    if (in.attr[loc] == nullptr || in.attr_size[loc] < size) { *err = ORC_ERR_INVALID; return nullptr; }   /* vertex.rs:148-150 panic */
    return in.attr[loc];
}

/* ---- colored_triangle.wgsl / hello_shader.wgsl (wgpu-cpu-tests/src/tests/colored_triangle.wgsl:15-37) ---- */
static void vs_colored_triangle(const VsIn& in, VsOut& out, const Resources&, int*) {
    const uint32_t vertex_index = in.vertex_index % 3u;
    const float x = (float)((int32_t)vertex_index - 1);
    const float y = (float)((int32_t)(vertex_index & 1u) * 2 - 1);
    out.position = {x, y, 0.0f, 1.0f};
    const float r = vertex_index == 0u ? 1.0f : 0.0f;
    const float g = vertex_index == 1u ? 1.0f : 0.0f;
    const float b = vertex_index == 2u ? 1.0f : 0.0f;
    store_vec4(out.inter + 0, {r, g, b, 1.0f});
}
static void fs_passthrough_color(const FsIn& in, FsOut& out, const Resources&, int*) {
    out.num_color = 1;
    out.color_location[0] = 0;
    out.color[0] = load_vec4(in.inter + 0);
}

/* ---- hello_mesh.wgsl (wgpu-cpu/examples/hello_mesh.wgsl:27-41) ---- */
static void vs_hello_mesh(const VsIn& in, VsOut& out, const Resources& res, int* err) {
    const uint8_t* cam = res.buffer(0, 0, 64, err);
    const uint8_t* p = need_attr(in, 0, 16, err);
    const uint8_t* c = need_attr(in, 1, 16, err);
    if (*err) return;
    float m[16];
    std::memcpy(m, cam, 64);
    out.position = mat4_mul_vec4(m, load_vec4(p));
    store_vec4(out.inter + 0, load_vec4(c));
}

/* ---- hello_texture.wgsl (wgpu-cpu/examples/hello_texture.wgsl:36-53) ---- */
static void vs_hello_texture(const VsIn& in, VsOut& out, const Resources& res, int* err) {
    const uint8_t* cam = res.buffer(0, 0, 64, err);
    const uint8_t* p = need_attr(in, 0, 16, err);
    const uint8_t* uv = need_attr(in, 1, 8, err);
    if (*err) return;
    float m[16];
    std::memcpy(m, cam, 64);
    out.position = mat4_mul_vec4(m, load_vec4(p));
    std::memcpy(out.inter + 0, uv, 8);
}
static void fs_hello_texture(const FsIn& in, FsOut& out, const Resources& res, int* err) {
    const float u = load_f32(in.inter + 0), v = load_f32(in.inter + 4);
    out.num_color = 1;
    out.color_location[0] = 0;
    out.color[0] = res.sample(1, 0, 1, 1, u, v, err);
}

/* ---- procedural.wgsl (wgpu-cpu_b200/shaders/procedural.wgsl, config C4) ---- */
static void vs_procedural(const VsIn& in, VsOut& out, const Resources&, int*) {
    const uint32_t vi = in.vertex_index % 6u;
    const float ux = (vi == 1u || vi == 4u || vi == 5u) ? 1.0f : 0.0f;
    const float uy = (vi == 2u || vi == 3u || vi == 5u) ? 1.0f : 0.0f;
    out.position = {ux * 2.0f - 1.0f, uy * 2.0f - 1.0f, 0.0f, 1.0f};
}
static void fs_procedural(const FsIn& in, FsOut& out, const Resources&, int*) {
    const float cx = in.position.x / 2560.0f - 2.0f;
    const float cy = in.position.y / 2160.0f - 1.0f;
    float zx = 0.0f, zy = 0.0f, acc = 0.0f;
    for (int i = 0; i < 64; i++) {
        const float nx = zx * zx - zy * zy + cx;
        const float ny = 2.0f * zx * zy + cy;
        const float r2 = nx * nx + ny * ny;
        const bool escaped = r2 > 4.0f;
        zx = escaped ? zx : nx;
        zy = escaped ? zy : ny;
        acc = acc + (escaped ? 0.0f : 1.0f);
    }
    const float t = acc / 64.0f;
    out.num_color = 1;
    out.color_location[0] = 0;
    out.color[0] = {t, t * t, 1.0f - t, 1.0f};
}

/* ---- features.wgsl ---- */
static void vs_features(const VsIn& in, VsOut& out, const Resources& res, int* err) {
    const uint8_t* prm = res.buffer(0, 0, 80, err);
    const uint8_t* p = need_attr(in, 0, 16, err);
    const uint8_t* c = need_attr(in, 1, 16, err);
    if (*err) return;
    float m[16];
    std::memcpy(m, prm, 64);
    const Vec4 off = load_vec4(prm + 64);
    const float shift = (float)in.instance_index;
    const Vec4 q = mat4_mul_vec4(m, load_vec4(p));
    out.position = {q.x + off.x * shift, q.y + off.y * shift, q.z + off.z * shift, q.w};
    store_vec4(out.inter + 0, load_vec4(c));
    store_u32(out.inter + 16, in.vertex_index + in.instance_index * 1000u);
}
static void fs_features(const FsIn& in, FsOut& out, const Resources&, int*) {
    /* as.rs: f32 -> u32 truncates and traps when out of range, so the shader bounds the
     * (possibly extrapolated) position with selects first */
    const float lx = in.position.x < 0.0f ? 0.0f : in.position.x;
    const float ly = in.position.y < 0.0f ? 0.0f : in.position.y;
    const uint32_t px = (uint32_t)(lx > 4096.0f ? 4096.0f : lx);
    const uint32_t py = (uint32_t)(ly > 4096.0f ? 4096.0f : ly);
    if ((px / 4u + py / 4u) % 3u == 0u) { out.killed = true; return; }
    const Vec4 color = load_vec4(in.inter + 0);
    const uint32_t tag = load_u32(in.inter + 16);
    const float t = (float)(tag % 7u) / 7.0f;
    const float facing = in.front_facing ? 1.0f : 0.0f;
    out.num_color = 1;
    out.color_location[0] = 0;
    out.color[0] = {color.x, color.y * t, color.z * facing, 1.0f};
}

/* ---- frag_depth.wgsl ---- */
static void vs_frag_depth(const VsIn& in, VsOut& out, const Resources& res, int* err) {
    const uint8_t* cam = res.buffer(0, 0, 64, err);
    const uint8_t* p = need_attr(in, 0, 16, err);
    const uint8_t* c = need_attr(in, 1, 16, err);
    if (*err) return;
    float m[16];
    std::memcpy(m, cam, 64);
    out.position = mat4_mul_vec4(m, load_vec4(p));
    store_vec4(out.inter + 0, load_vec4(c));
}
static void fs_frag_depth(const FsIn& in, FsOut& out, const Resources&, int*) {
    const Vec4 color = load_vec4(in.inter + 0);
    out.has_frag_depth = true;
    out.frag_depth = 1.0f - in.position.z * color.x;
    out.num_color = 1;
    out.color_location[0] = 0;
    out.color[0] = color;
}

/* ---- early_force.wgsl / early_allow.wgsl (same stages; they differ in @early_depth_test) ---- */
static void fs_early_depth(const FsIn& in, FsOut& out, const Resources&, int*) {
    const Vec4 color = load_vec4(in.inter + 0);
    if (color.y > 0.6f) { out.killed = true; return; }
    out.has_frag_depth = true;
    out.frag_depth = in.position.z * 0.5f;
    out.num_color = 1;
    out.color_location[0] = 0;
    out.color[0] = color;
}

/* ---- mrt.wgsl: outputs in declaration order (locations 0, 2, 1) ---- */
static void fs_mrt(const FsIn& in, FsOut& out, const Resources&, int*) {
    const Vec4 tint = load_vec4(in.inter + 0);
    out.num_color = 3;
    out.color_location[0] = 0; out.color[0] = tint;
    out.color_location[1] = 2; out.color[1] = {in.position.z, tint.x * tint.y, 0.25f, 1.0f};
    out.color_location[2] = 1; out.color[2] = {tint.z, tint.y, tint.x, 0.5f};
}

/* ---- depth_only.wgsl: no inter-stage values, no outputs ---- */
static void vs_depth_only(const VsIn& in, VsOut& out, const Resources& res, int* err) {
    const uint8_t* cam = res.buffer(0, 0, 64, err);
    const uint8_t* p = need_attr(in, 0, 16, err);
    if (*err) return;
    float m[16];
    std::memcpy(m, cam, 64);
    out.position = mat4_mul_vec4(m, load_vec4(p));
}
static void fs_depth_only(const FsIn&, FsOut& out, const Resources&, int*) { out.num_color = 0; }

/* ---- prim_index.wgsl ---- */
static void fs_prim_index(const FsIn& in, FsOut& out, const Resources&, int*) {
    const float low = (float)(in.primitive_index & 255u) / 255.0f;
    const float high = (float)((in.primitive_index >> 8) & 255u) / 255.0f;
    out.num_color = 1;
    out.color_location[0] = 0;
    out.color[0] = {low, high, (float)in.sample_index + (float)(in.sample_mask & 1u) * 0.5f, 1.0f};
}

/* ---- perspective.wgsl: inter-stage layout vec4 @0, vec2 @16, f32 @24 (bindings.rs:307-346) ---- */
static void vs_perspective(const VsIn& in, VsOut& out, const Resources& res, int* err) {
    const uint8_t* cam = res.buffer(0, 0, 64, err);
    const uint8_t* p = need_attr(in, 0, 16, err);
    const uint8_t* c = need_attr(in, 1, 16, err);
    if (*err) return;
    float m[16];
    std::memcpy(m, cam, 64);
    out.position = mat4_mul_vec4(m, load_vec4(p));
    const Vec4 tint = load_vec4(c);
    store_vec4(out.inter + 0, tint);
    std::memcpy(out.inter + 16, &tint.x, 4);
    std::memcpy(out.inter + 20, &tint.y, 4);
    std::memcpy(out.inter + 24, &tint.z, 4);
}
static void fs_perspective(const FsIn& in, FsOut& out, const Resources&, int*) {
    const Vec4 corrected = load_vec4(in.inter + 0);
    float sx, sy, flat;
    std::memcpy(&sx, in.inter + 16, 4);
    std::memcpy(&sy, in.inter + 20, 4);
    std::memcpy(&flat, in.inter + 24, 4);
    out.num_color = 1;
    out.color_location[0] = 0;
    out.color[0] = {corrected.x, corrected.y * 0.5f + sy * 0.5f, sx, flat};
}

static const ShaderInfo SHADERS[ORC_SHADER_COUNT] = {
    /* colored_triangle */ {vs_colored_triangle, fs_passthrough_color, 1, {{0, 0, 4, VAR_F32, INTERP_LINEAR}}, 0},
    /* hello_mesh */       {vs_hello_mesh, fs_passthrough_color, 1, {{0, 0, 4, VAR_F32, INTERP_LINEAR}}, 0},
    /* hello_texture */    {vs_hello_texture, fs_hello_texture, 1, {{0, 0, 2, VAR_F32, INTERP_LINEAR}}, 0},
    /* procedural */       {vs_procedural, fs_procedural, 0, {}, 0},
    /* features */         {vs_features, fs_features, 2, {{0, 0, 4, VAR_F32, INTERP_LINEAR}, {1, 16, 1, VAR_U32, INTERP_FLAT}}, 0},
    /* frag_depth */       {vs_frag_depth, fs_frag_depth, 1, {{0, 0, 4, VAR_F32, INTERP_LINEAR}}, 0},
    /* early_force */      {vs_frag_depth, fs_early_depth, 1, {{0, 0, 4, VAR_F32, INTERP_LINEAR}}, 1},
    /* early_allow */      {vs_frag_depth, fs_early_depth, 1, {{0, 0, 4, VAR_F32, INTERP_LINEAR}}, 2},
    /* mrt */              {vs_frag_depth, fs_mrt, 1, {{0, 0, 4, VAR_F32, INTERP_LINEAR}}, 0},
    /* depth_only */       {vs_depth_only, fs_depth_only, 0, {}, 1},
    /* prim_index */       {vs_depth_only, fs_prim_index, 0, {}, 0},
    /* perspective */      {vs_perspective, fs_perspective, 3, {{0, 0, 4, VAR_F32, INTERP_PERSPECTIVE}, {1, 16, 2, VAR_F32, INTERP_LINEAR},
                                                                 {2, 24, 1, VAR_F32, INTERP_FLAT}}, 0},
};

const ShaderInfo* shader_info(uint32_t shader) {
    return shader < ORC_SHADER_COUNT ? &SHADERS[shader] : nullptr;
}

}  // namespace orc

extern "C" void orc_mat4_mul_vec4(const float m[16], const float v[4], float out[4]) {
    const orc::Vec4 r = orc::mat4_mul_vec4(m, {v[0], v[1], v[2], v[3]});
    out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}
