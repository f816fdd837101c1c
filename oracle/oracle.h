/*
 * oracle.h -- CPU restatement of wgpu-cpu's render-pass draw path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (wgpu_b200/) may include,
 * link or call this.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker or the
 * timed CPU baseline.
 *
 * The reference (jgraef/wgpu-cpu) is Rust and cannot be compiled in this image
 * (no rustc/cargo; un-vendored git dependencies), so this is a behavioural
 * restatement in plain C++ of the files listed in SURVEY.md section 8(c).  Every
 * function cites the reference file:line it follows (paths relative to
 * /root/reference/).
 *
 * Parity pinning: the unit functions below are checked against every exact-value
 * known-answer test the reference holds for this path (tests/test_oracle_kat.py):
 * Bresenham point lists, Cohen-Sutherland end points, triangle-clip outputs,
 * strip order + winding, index resolution, texture layout, sort triples,
 * mat4*vec4 = [14,19,16,21], the vertex-index triangle positions.  `to_raster`'s
 * divide/scale/truncate, the shoelace barycentrics, `texel_coordinate` rounding
 * and `f32_to_u8` truncation have NO test in the reference and its golden PNGs
 * are git-LFS pointers: for those steps parity is UNPINNED -- the oracle follows
 * the source text and IEEE-754 semantics (built with -ffp-contract=off).
 */
#ifndef WGPU_CPU_ORACLE_H
#define WGPU_CPU_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* enumerations: numeric values shared with include/wgpu_b200.h by convention */
enum { ORC_TOPO_POINT_LIST = 0, ORC_TOPO_LINE_LIST = 1, ORC_TOPO_LINE_STRIP = 2,
       ORC_TOPO_TRIANGLE_LIST = 3, ORC_TOPO_TRIANGLE_STRIP = 4 };
enum { ORC_INDEX_NONE = 0, ORC_INDEX_U16 = 1, ORC_INDEX_U32 = 2 };
enum { ORC_FRONT_CCW = 0, ORC_FRONT_CW = 1 };
enum { ORC_CULL_NONE = 0, ORC_CULL_FRONT = 1, ORC_CULL_BACK = 2 };
enum { ORC_CMP_NEVER = 1, ORC_CMP_LESS = 2, ORC_CMP_EQUAL = 3, ORC_CMP_LESS_EQUAL = 4,
       ORC_CMP_GREATER = 5, ORC_CMP_NOT_EQUAL = 6, ORC_CMP_GREATER_EQUAL = 7, ORC_CMP_ALWAYS = 8 };
enum { ORC_FMT_RGBA8_UNORM = 0, ORC_FMT_RGBA8_UNORM_SRGB = 1, ORC_FMT_BGRA8_UNORM = 2,
       ORC_FMT_BGRA8_UNORM_SRGB = 3, ORC_FMT_R8_UNORM = 4, ORC_FMT_RG8_UNORM = 5,
       ORC_FMT_RGBA8_SNORM = 6, ORC_FMT_DEPTH32_FLOAT = 7 };
enum { ORC_ADDR_CLAMP_TO_EDGE = 0, ORC_ADDR_REPEAT = 1, ORC_ADDR_MIRROR_REPEAT = 2 };
enum { ORC_STEP_VERTEX = 0, ORC_STEP_INSTANCE = 1 };
enum { ORC_BIND_NONE = 0, ORC_BIND_BUFFER = 1, ORC_BIND_TEXTURE = 2, ORC_BIND_SAMPLER = 3 };

/* built-in shaders (hand-written restatements of the WGSL in the op order the
 * reference JIT emits) */
enum { ORC_SHADER_COLORED_TRIANGLE = 0, /* colored_triangle.wgsl == hello_shader.wgsl */
       ORC_SHADER_HELLO_MESH = 1, ORC_SHADER_HELLO_TEXTURE = 2,
       ORC_SHADER_PROCEDURAL = 3,       /* wgpu_b200/shaders/procedural.wgsl (config C4) */
       ORC_SHADER_FEATURES = 4,         /* wgpu_b200/shaders/features.wgsl: flat varying, discard, instance_index */
       ORC_SHADER_FRAG_DEPTH = 5,       /* wgpu_b200/shaders/frag_depth.wgsl */
       ORC_SHADER_EARLY_FORCE = 6,      /* wgpu_b200/shaders/early_force.wgsl: @early_depth_test(force) + discard + frag_depth */
       ORC_SHADER_EARLY_ALLOW = 7,      /* wgpu_b200/shaders/early_allow.wgsl: EarlyDepthTest::Allow (early and late test) */
       ORC_SHADER_MRT = 8,              /* wgpu_b200/shaders/mrt.wgsl: three colour attachments */
       ORC_SHADER_DEPTH_ONLY = 9,       /* wgpu_b200/shaders/depth_only.wgsl: no fragment outputs, @early_depth_test(force) */
       ORC_SHADER_PRIM_INDEX = 10,      /* wgpu_b200/shaders/prim_index.wgsl: primitive_index, sample_index, sample_mask */
       ORC_SHADER_PERSPECTIVE = 11,     /* wgpu_b200/shaders/perspective.wgsl: default (perspective-correct) interpolation -- not in the reference */
       ORC_SHADER_COUNT = 12 };

enum { ORC_OK = 0, ORC_ERR_INVALID = 1, ORC_ERR_OUT_OF_BOUNDS = 2, ORC_ERR_UNSUPPORTED = 3,
       ORC_ERR_W_ZERO = 4 };

#define ORC_MAX_GROUPS 4
#define ORC_MAX_BINDINGS 4
#define ORC_MAX_VERTEX_BUFFERS 8
#define ORC_MAX_ATTRS 16
#define ORC_MAX_COLOR 4

typedef struct { uint8_t* data; uint32_t width, height, format; } orc_texture;
typedef struct { uint32_t address_u, address_v; } orc_sampler;
typedef struct { const uint8_t* data; uint64_t size; } orc_buffer;
typedef struct {
    uint32_t kind;
    orc_buffer buffer;
    orc_texture texture;
    orc_sampler sampler;
} orc_binding;
typedef struct { orc_binding b[ORC_MAX_GROUPS][ORC_MAX_BINDINGS]; } orc_bindings;

typedef struct { uint32_t location, buffer, offset, size; } orc_vertex_attr;
typedef struct { uint32_t stride, step_mode; } orc_vertex_buffer_layout;

typedef struct {
    uint32_t shader;
    uint32_t topology, strip_index_format, front_face, cull_mode;
    uint32_t has_depth_state, depth_compare, depth_write;
    uint32_t has_fragment;
    uint32_t num_vertex_buffers;
    orc_vertex_buffer_layout vb[ORC_MAX_VERTEX_BUFFERS];
    uint32_t num_attrs;
    orc_vertex_attr attrs[ORC_MAX_ATTRS];
} orc_pipeline;

typedef struct {
    uint32_t num_color;
    orc_texture color[ORC_MAX_COLOR];
    uint32_t color_clear[ORC_MAX_COLOR];      /* 1 = LoadOp::Clear */
    double clear_color[ORC_MAX_COLOR][4];     /* wgpu::Color is f64 */
    uint32_t has_depth;
    orc_texture depth;
    uint32_t depth_clear;
    float clear_depth;
    uint32_t ext_features;                    /* ORC_EXT_SRGB_ENCODE applies to the clear colour too */
} orc_pass;

typedef struct {
    float vp_x, vp_y, vp_w, vp_h, vp_min_depth, vp_max_depth;
    uint32_t sc_x, sc_y, sc_w, sc_h;
    /* Behaviour BEYOND the reference (it accepts and ignores these states): the product's opt-in WGB_FEATURE_* bits,
     * restated from the WebGPU specification -- there is nothing in the reference to pin them against.  0 = the
     * reference's behaviour. */
    uint32_t ext_features;
    uint32_t color_write_mask[ORC_MAX_COLOR]; /* ORC_EXT_COLOR_WRITE_MASK: R=1 G=2 B=4 A=8 */
    /* ORC_EXT_BLEND: per target {enabled, colour src factor, dst factor, operation, alpha src, dst, operation}
     * (wgpu::BlendFactor / BlendOperation numbering of include/wgpu_b200.h) and the blend constant */
    uint32_t blend[ORC_MAX_COLOR][7];
    float blend_constant[4];
} orc_raster_state;
#define ORC_EXT_VIEWPORT_DEPTH_RANGE 1u   /* fragment depth = min_depth + ndc.z * (max_depth - min_depth) */
#define ORC_EXT_COLOR_WRITE_MASK 2u
#define ORC_EXT_SRGB_ENCODE 4u            /* *Srgb targets store the sRGB-encoded value */
#define ORC_EXT_BLEND 16u                 /* WebGPU blend equation on [0,1] values against the stored (8-bit) texel */

typedef struct {
    uint32_t indexed;
    uint32_t first, count;           /* first index / first vertex, count */
    int32_t base_vertex;
    uint32_t first_instance, instance_count;
    uint32_t index_format;
    orc_buffer index_buffer;
    orc_buffer vertex_buffers[ORC_MAX_VERTEX_BUFFERS];
} orc_draw;

typedef struct {
    uint64_t vertices_processed;     /* state.rs:517 */
    uint64_t primitives_assembled;
    uint64_t primitives_culled;
    uint64_t primitives_drawn;       /* post-clip, state.rs:516 */
    uint64_t fragments_shaded;       /* fragment shader invocations */
    uint64_t fragments_written;      /* passed the depth test and wrote colour */
} orc_stats;

/* ---- the path ---- */
/* State::load: LoadOp::Clear of every attachment (state.rs:135-145) */
int orc_pass_load(const orc_pass* pass);
/* DrawCall::execute + draw<> (state.rs:238-593).  `coverage`, if non-null, is a
 * width*height u32 array incremented once per rasterised fragment that reached
 * the fragment stage (after the scissor test).  */
int orc_draw_execute(const orc_pass* pass, const orc_pipeline* pipe, const orc_raster_state* rs,
                     const orc_bindings* bindings, const orc_draw* draw, orc_stats* stats,
                     uint32_t* coverage);
/* default viewport / scissor for a framebuffer (state.rs:604-628) */
void orc_default_raster_state(uint32_t width, uint32_t height, orc_raster_state* out);

/* ---- unit functions, exposed for the reference's known-answer tests ---- */
void orc_bubblesort3_i64(int64_t v[3]);                                   /* util/sort.rs:3-22 */
/* util/scanline.rs:14-198: rows as (y, x1, x2) triples; returns row count */
int orc_scanlines(const uint32_t tri[6], uint32_t* rows, int max_rows);
/* render_pass/clipper.rs:625-802: in = 3 vec4; out = n*(3 vec4) and n*(3 vec3 barycentrics) */
int orc_clip_triangle(const float in[12], float* out_pos, float* out_bary, int max_tris);
/* render_pass/clipper.rs:268-356: returns 1 if a segment survives */
int orc_clip_line(const float in[8], float out_pos[8], float out_alpha[2]);
/* util/bresenham.rs:78-153: points as (x,y) pairs + t; returns count */
int orc_bresenham(const uint32_t start[2], const uint32_t end[2], uint32_t* pts, float* ts, int max_pts);
/* render_pass/primitive.rs:360-487: triangle strip assembly over 0..n (separator = 0xFFFFFFFF if sep) */
int orc_tri_strip(const uint32_t* items, int n, int separated, uint32_t* out_tris, int max_tris);
/* render_pass/primitive.rs:169-197: 1 = Ccw, 0 = Cw */
int orc_front_face_ccw(const float clip[12]);
/* render_pass/index.rs:45-88 */
int orc_resolve_index(const uint8_t* index_data, uint64_t size, uint32_t format, int32_t base_vertex,
                      uint32_t i, int separated, uint32_t* out_vertex, int* is_separator);
/* render_pass/raster.rs:129-160 */
int orc_to_raster(const orc_raster_state* rs, const float clip[4], uint32_t fb[2], float frag[4]);
uint32_t orc_texel_coordinate(float x, uint32_t address_mode, uint32_t size);   /* binding.rs:151-164 */
uint8_t orc_f32_to_u8(float v);                                               /* texture.rs:376-379 */
uint64_t orc_texture_byte_size(uint32_t format, uint32_t w, uint32_t h, uint32_t layers); /* texture.rs:257-273 */
uint64_t orc_texel_byte_offset(uint32_t format, uint32_t w, uint32_t h, uint32_t x, uint32_t y, uint32_t z);
/* naga-cranelift binary.rs:297-323: column-major mat4 * vec4, no FMA */
void orc_mat4_mul_vec4(const float m[16], const float v[4], float out[4]);
/* run a built-in vertex shader once (for the vertex-index triangle KAT, naga-cranelift tests.rs:47-110) */
int orc_run_vertex_shader(uint32_t shader, uint32_t vertex_index, uint32_t instance_index,
                          const orc_bindings* bindings, float out_position[4], float out_varyings[16]);

#ifdef __cplusplus
}
#endif
#endif
