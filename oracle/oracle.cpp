/*
 * oracle.cpp -- CPU restatement of wgpu-cpu's render-pass draw path.
 * TEST INFRASTRUCTURE ONLY (see oracle.h for the rules and the parity-pinning note).
 *
 * Build: g++ -O2 -ffp-contract=off (no fast-math): Rust never contracts a*b+c into
 * an FMA and every f32 op below must round exactly like the reference's.
 * All paths cited are relative to /root/reference/.
 */
#include "oracle_internal.h"

#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <deque>
#include <vector>

namespace orc {

/* ------------------------------------------------------------------------- */
/* small vector helpers: one IEEE op per call, no contraction                  */
/* ------------------------------------------------------------------------- */
static inline Vec4 v4_scale(Vec4 a, float s) { return {a.x * s, a.y * s, a.z * s, a.w * s}; }
static inline Vec4 v4_add(Vec4 a, Vec4 b) { return {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }

struct Bary3 { float c[3]; };
static inline Bary3 b3_scale(Bary3 a, float s) { return {{a.c[0] * s, a.c[1] * s, a.c[2] * s}}; }
static inline Bary3 b3_add(Bary3 a, Bary3 b) { return {{a.c[0] + b.c[0], a.c[1] + b.c[1], a.c[2] + b.c[2]}}; }

/* Rust `as u32` on f32: saturating, NaN -> 0 (nalgebra try_cast::<u32> goes through
 * simba's SubsetOf<f32> for u32, whose is_in_subset is always true and whose
 * conversion is `as`; raster.rs:153) */
static inline uint32_t f32_as_u32(float v) {
    if (!(v == v)) return 0u;
    if (v <= 0.0f) return 0u;
    if (v >= 4294967296.0f) return 0xFFFFFFFFu;
    return (uint32_t)v;
}

/* ------------------------------------------------------------------------- */
/* util/sort.rs:3-22 -- CAS network (0,1),(0,2),(1,2), swap only if Greater    */
/* ------------------------------------------------------------------------- */
struct P2i { int64_t x, y; };
template <class T, class Gt>
static inline void bubblesort3(T v[3], Gt greater) {
    if (greater(v[0], v[1])) { T t = v[0]; v[0] = v[1]; v[1] = t; }
    if (greater(v[0], v[2])) { T t = v[0]; v[0] = v[2]; v[2] = t; }
    if (greater(v[1], v[2])) { T t = v[1]; v[1] = v[2]; v[2] = t; }
}

/* ------------------------------------------------------------------------- */
/* util/scanline.rs:14-198                                                     */
/* ------------------------------------------------------------------------- */
struct Row { int64_t y, x1, x2; };
/* emits rows top->bottom with x1 <= x2 (Scanline::new_from_isize swaps, :155-170) */
template <class Emit>
static inline void scanlines(const uint32_t fb[3][2], Emit emit) {
    P2i tri[3];
    for (int i = 0; i < 3; i++) { tri[i].x = (int64_t)fb[i][0]; tri[i].y = (int64_t)fb[i][1]; }
    bubblesort3(tri, [](const P2i& a, const P2i& b) { return a.y > b.y; });   /* :20 */
    const P2i a = tri[0], b = tri[1], c = tri[2];
    const P2i ac = {c.x - a.x, c.y - a.y};
    const P2i ab = {b.x - a.x, b.y - a.y};
    const P2i bc = {c.x - b.x, c.y - b.y};

    auto put = [&](int64_t y, int64_t x1, int64_t x2) {
        if (x1 > x2) { int64_t t = x1; x1 = x2; x2 = t; }
        emit(Row{y, x1, x2});
    };

    if (ac.y == 0) {                                  /* :28-35 DegenerateLine */
        bubblesort3(tri, [](const P2i& p, const P2i& q) { return p.x > q.x; });
        put(a.y, tri[0].x, tri[2].x);
        return;
    }
    int64_t y = a.y, i = 0;
    if (ab.y != 0) {                                  /* State::First :104-128 */
        while (y < b.y) {
            const int64_t x1 = a.x + (ac.x * i) / ac.y;   /* Rust `/` truncates toward zero, as C++ */
            const int64_t x2 = a.x + (ab.x * i) / ab.y;
            put(y, x1, x2);
            y += 1; i += 1;
        }
        if (bc.y == 0) return;                        /* :120-122: flat-bottom loses row c.y */
    }
    int64_t i2 = 0;                                   /* State::Second :129-146 */
    while (y <= c.y) {
        const int64_t x1 = a.x + (ac.x * i) / ac.y;
        const int64_t x2 = b.x + (bc.x * i2) / bc.y;
        i2 += 1;
        put(y, x1, x2);
        y += 1; i += 1;
    }
}

/* ------------------------------------------------------------------------- */
/* render_pass/clipper.rs:133-205 -- ClipVolume::WEBGPU                        */
/* ------------------------------------------------------------------------- */
static const float CLIP_PLANES[6][4] = {
    {1.0f, 0.0f, 0.0f, 1.0f},    /* left:   x < -w */
    {-1.0f, 0.0f, 0.0f, 1.0f},   /* right:  x >  w */
    {0.0f, 1.0f, 0.0f, 1.0f},    /* bottom: y < -w */
    {0.0f, -1.0f, 0.0f, 1.0f},   /* top:    y >  w */
    {0.0f, 0.0f, 1.0f, 0.0f},    /* front:  z <  0 */
    {0.0f, 0.0f, -1.0f, 1.0f},   /* back:   z >  w */
};
/* ClipPlane::clip_distance = point.dot(plane) (clipper.rs:140-143).  nalgebra 0.34
 * (not vendored) special-cases 4-vectors in Matrix::dotx (base/blas.rs):
 *   a = p0*n0; b = p1*n1; c = p2*n2; d = p3*n3; a += c; b += d; return a + b  */
static inline float clip_distance(int plane, Vec4 p) {
    const float* n = CLIP_PLANES[plane];
    float a = p.x * n[0];
    float b = p.y * n[1];
    const float c = p.z * n[2];
    const float d = p.w * n[3];
    a += c;
    b += d;
    return a + b;
}

/* clipper.rs:583-612 tri_clip::Vertex and util/interpolation.rs:84-89 lerp */
struct ClipVertex { Vec4 position; Bary3 barycentric; };
static inline ClipVertex cv_lerp(const ClipVertex& x0, const ClipVertex& x1, float t) {
    const float s = 1.0f - t;
    ClipVertex r;
    r.position = v4_add(v4_scale(x0.position, s), v4_scale(x1.position, t));
    r.barycentric = b3_add(b3_scale(x0.barycentric, s), b3_scale(x1.barycentric, t));
    return r;
}
struct ClipTri { ClipVertex v[3]; };

/* clipper.rs:614-729 clip_tri_against_plane: the 8-case table, with the exact edge
 * directions the reference uses */
static inline int clip_tri_against_plane(const ClipTri& tri, int plane, ClipTri out[2]) {
    float d[3];
    bool outside[3];
    for (int i = 0; i < 3; i++) { d[i] = clip_distance(plane, tri.v[i].position); outside[i] = d[i] < 0.0f; }
    auto clip = [&](int i, int j) {
        const float t = d[i] / (d[i] - d[j]);
        return cv_lerp(tri.v[i], tri.v[j], t);
    };
    const int code = (outside[0] ? 1 : 0) | (outside[1] ? 2 : 0) | (outside[2] ? 4 : 0);
    switch (code) {
        case 7: return 0;                                            /* [t,t,t] */
        case 0: out[0] = tri; return 1;                              /* [f,f,f] */
        case 3: {                                                    /* [t,t,f] :640-652 */
            ClipVertex a_new = clip(2, 0), b_new = clip(1, 2);
            out[0] = ClipTri{{a_new, b_new, tri.v[2]}};
            return 1;
        }
        case 6: {                                                    /* [f,t,t] :653-665 */
            ClipVertex b_new = clip(0, 1), c_new = clip(2, 0);
            out[0] = ClipTri{{tri.v[0], b_new, c_new}};
            return 1;
        }
        case 5: {                                                    /* [t,f,t] :666-678 */
            ClipVertex a_new = clip(0, 1), c_new = clip(1, 2);
            out[0] = ClipTri{{a_new, tri.v[1], c_new}};
            return 1;
        }
        case 1: {                                                    /* [t,f,f] :679-694 */
            ClipVertex b_new = clip(0, 1), c_new = clip(0, 2);
            out[0] = ClipTri{{b_new, tri.v[1], c_new}};
            out[1] = ClipTri{{tri.v[1], tri.v[2], c_new}};
            return 2;
        }
        case 2: {                                                    /* [f,t,f] :695-710 */
            ClipVertex a_new = clip(0, 1), c_new = clip(1, 2);
            out[0] = ClipTri{{tri.v[0], a_new, c_new}};
            out[1] = ClipTri{{tri.v[0], c_new, tri.v[2]}};
            return 2;
        }
        default: {                                                   /* 4: [f,f,t] :711-726 */
            ClipVertex b_new = clip(1, 2), a_new = clip(0, 2);
            out[0] = ClipTri{{tri.v[0], tri.v[1], b_new}};
            out[1] = ClipTri{{tri.v[0], b_new, a_new}};
            return 2;
        }
    }
}

/* clipper.rs:756-779 TriClipper::clip_tri: FIFO queue, plane by plane */
static void clip_tri(const Vec4 pos[3], std::deque<ClipTri>& queue) {
    queue.clear();
    ClipTri t0;
    for (int i = 0; i < 3; i++) {
        t0.v[i].position = pos[i];
        t0.v[i].barycentric = Bary3{{0.0f, 0.0f, 0.0f}};
        t0.v[i].barycentric.c[i] = 1.0f;               /* Barycentric::at_vertex(i) */
    }
    queue.push_back(t0);
    for (int plane = 0; plane < 6; plane++) {
        const size_t n = queue.size();
        for (size_t k = 0; k < n; k++) {
            ClipTri tri = queue.front();
            queue.pop_front();
            ClipTri out[2];
            const int m = clip_tri_against_plane(tri, plane, out);
            for (int j = 0; j < m; j++) queue.push_back(out[j]);
        }
    }
}

/* clipper.rs:268-356 cohen_sutherland::clip */
static bool clip_line(Vec4 points[2], float alphas_out[2]) {
    float cd[2][6];
    unsigned outcodes[2] = {0, 0};
    for (int p = 0; p < 2; p++)
        for (int i = 0; i < 6; i++) {
            cd[p][i] = clip_distance(i, points[p]);
            if (cd[p][i] < 0.0f) outcodes[p] |= 1u << i;
        }
    const unsigned o_or = outcodes[0] | outcodes[1];
    const unsigned o_and = outcodes[0] & outcodes[1];
    if (o_or == 0) { alphas_out[0] = 0.0f; alphas_out[1] = 1.0f; return true; }
    if (o_and != 0) return false;
    float alphas[2] = {0.0f, 1.0f};
    for (int plane = 0; plane < 6; plane++) {
        const unsigned bit = 1u << plane;
        if (o_or & bit) {
            const float a = cd[0][plane], b = cd[1][plane];
            const float alpha = a / (a - b);
            if (outcodes[0] & bit) { if (alpha > alphas[0]) alphas[0] = alpha; }
            else { if (alpha < alphas[1]) alphas[1] = alpha; }
            if (alphas[0] > alphas[1]) return false;
        }
    }
    const Vec4 p0 = points[0], p1 = points[1];
    /* update_point(i): lerp(points[0], points[1], alphas[i]) -- note that when both
     * points are updated the second lerp reads the ALREADY UPDATED points[0]
     * (the closure mutates `points` in place, clipper.rs:333-346) */
    Vec4 cur0 = p0;
    if (outcodes[0] != 0) {
        const float t = alphas[0];
        cur0 = v4_add(v4_scale(p0, 1.0f - t), v4_scale(p1, t));
        points[0] = cur0;
    }
    if (outcodes[1] != 0) {
        const float t = alphas[1];
        points[1] = v4_add(v4_scale(cur0, 1.0f - t), v4_scale(p1, t));
    }
    alphas_out[0] = alphas[0];
    alphas_out[1] = alphas[1];
    return true;
}

/* ------------------------------------------------------------------------- */
/* util/bresenham.rs:17-153                                                    */
/* ------------------------------------------------------------------------- */
struct Bresenham {
    P2i x, d;
    int64_t x1, diff;
    int octant;
    float t, dt;

    static int octant_from_points(P2i start, P2i end) {                /* :22-46 */
        P2i d = {end.x - start.x, end.y - start.y};
        int octant = 0;
        if (d.y < 0) { d.x = -d.x; d.y = -d.y; octant += 4; }
        if (d.x < 0) { int64_t tmp = d.x; d.x = d.y; d.y = -tmp; octant += 2; }
        if (d.x < d.y) octant += 1;
        return octant;
    }
    static P2i to_octant0(int o, P2i p) {                              /* :48-62 */
        switch (o) {
            case 0: return {p.x, p.y};   case 1: return {p.y, p.x};
            case 2: return {p.y, -p.x};  case 3: return {-p.x, p.y};
            case 4: return {-p.x, -p.y}; case 5: return {-p.y, -p.x};
            case 6: return {-p.y, p.x};  default: return {p.x, -p.y};
        }
    }
    static P2i from_octant0(int o, P2i p) {                            /* :64-78 */
        switch (o) {
            case 0: return {p.x, p.y};   case 1: return {p.y, p.x};
            case 2: return {-p.y, p.x};  case 3: return {-p.x, p.y};
            case 4: return {-p.x, -p.y}; case 5: return {-p.y, -p.x};
            case 6: return {p.y, -p.x};  default: return {p.x, -p.y};
        }
    }
    Bresenham(const uint32_t s[2], const uint32_t e[2]) {              /* :103-127 */
        P2i start = {(int64_t)s[0], (int64_t)s[1]}, end = {(int64_t)e[0], (int64_t)e[1]};
        octant = octant_from_points(start, end);
        start = to_octant0(octant, start);
        end = to_octant0(octant, end);
        d = {end.x - start.x, end.y - start.y};
        dt = 1.0f / (float)d.x;
        x = start; x1 = end.x; diff = d.y - d.x; t = 0.0f;
    }
    bool next(uint32_t pt[2], float* t_out) {                          /* :130-153 */
        if (x.x >= x1) return false;
        const P2i p = x;
        if (diff >= 0) { x.y += 1; diff -= d.x; }
        diff += d.y;
        const float tt = t;
        t += dt;
        x.x += 1;
        const P2i q = from_octant0(octant, p);
        pt[0] = (uint32_t)q.x; pt[1] = (uint32_t)q.y;
        *t_out = tt;
        return true;
    }
};

/* ------------------------------------------------------------------------- */
/* render_pass/raster.rs:122-168 ToRaster                                      */
/* ------------------------------------------------------------------------- */
struct ToRaster {
    float tx, ty, sx, sy;
    uint32_t sc0x, sc0y, sc1x, sc1y;
    bool depth_range; float depth_min, depth_scale;     /* ORC_EXT_VIEWPORT_DEPTH_RANGE (not in the reference) */
    explicit ToRaster(const orc_raster_state& rs) {
        depth_range = (rs.ext_features & ORC_EXT_VIEWPORT_DEPTH_RANGE) != 0;
        depth_min = rs.vp_min_depth; depth_scale = rs.vp_max_depth - rs.vp_min_depth;
        tx = rs.vp_x + 0.5f * rs.vp_w;                /* :131 */
        ty = rs.vp_y + 0.5f * rs.vp_h;
        sx = 0.5f * rs.vp_w;                          /* :132-133 */
        sy = 0.5f * rs.vp_h;
        sy *= -1.0f;
        sc0x = rs.sc_x; sc0y = rs.sc_y;
        sc1x = rs.sc_x + rs.sc_w; sc1y = rs.sc_y + rs.sc_h;
    }
    /* :145-160; returns false on w == 0 (the reference panics "w=0") */
    bool to_raster(Vec4 clip, uint32_t fb[2], Vec4& frag) const {
        if (clip.w == 0.0f) return false;
        /* Point3::from_homogeneous: coords / w, a true division per component */
        const float nx = clip.x / clip.w, ny = clip.y / clip.w;
        float nz = clip.z / clip.w;
        if (depth_range) nz = depth_min + nz * depth_scale;
        const float pd = 1.0f / clip.w;
        float vx = nx * sx + tx;
        float vy = ny * sy + ty;
        vx = fmaxf(vx, 0.0f);
        vy = fmaxf(vy, 0.0f);
        fb[0] = f32_as_u32(vx);
        fb[1] = f32_as_u32(vy);
        frag = {vx, vy, nz, pd};
        return true;
    }
    bool in_scissor(uint32_t x, uint32_t y) const {   /* :162-167 */
        return x >= sc0x && x < sc1x && y >= sc0y && y < sc1y;
    }
};

/* ------------------------------------------------------------------------- */
/* texture.rs: layout, texel encode/decode, clears                             */
/* ------------------------------------------------------------------------- */
static int bytes_per_texel(uint32_t format) {          /* texture.rs:457-509 (subset used on the path) */
    switch (format) {
        case ORC_FMT_R8_UNORM: return 1;
        case ORC_FMT_RG8_UNORM: return 2;
        case ORC_FMT_RGBA8_UNORM: case ORC_FMT_RGBA8_UNORM_SRGB: case ORC_FMT_BGRA8_UNORM:
        case ORC_FMT_BGRA8_UNORM_SRGB: case ORC_FMT_RGBA8_SNORM: case ORC_FMT_DEPTH32_FLOAT: return 4;
        default: return 0;
    }
}
static inline uint8_t f32_to_u8(float value) {         /* texture.rs:376-379 */
    float v = value * 255.0f;
    /* f32::clamp(0, MAX): NaN stays NaN, then `as u8` saturates (NaN -> 0), truncating */
    if (!(v == v)) return 0;
    if (v < 0.0f) v = 0.0f;
    if (v > 255.0f) v = 255.0f;
    return (uint8_t)v;
}
static inline float srgb_oetf(float v) {                /* ORC_EXT_SRGB_ENCODE (not in the reference) */
    if (!(v > 0.0031308f)) return v * 12.92f;
    return 1.055f * powf(v, 1.0f / 2.4f) - 0.055f;
}
/* TexelWriter::from_color (texture.rs:373-411); returns texel size or 0 if the format panics */
static int encode_color(Vec4 c, uint32_t format, uint8_t out[4], bool srgb_encode = false) {
    if (srgb_encode && (format == ORC_FMT_RGBA8_UNORM_SRGB || format == ORC_FMT_BGRA8_UNORM_SRGB)) {
        c.x = srgb_oetf(c.x); c.y = srgb_oetf(c.y); c.z = srgb_oetf(c.z);
    }
    switch (format) {
        case ORC_FMT_R8_UNORM: out[0] = f32_to_u8(c.x); return 1;
        case ORC_FMT_RG8_UNORM: out[0] = f32_to_u8(c.x); out[1] = f32_to_u8(c.y); return 2;
        case ORC_FMT_RGBA8_UNORM: case ORC_FMT_RGBA8_UNORM_SRGB: case ORC_FMT_RGBA8_SNORM:
            out[0] = f32_to_u8(c.x); out[1] = f32_to_u8(c.y); out[2] = f32_to_u8(c.z); out[3] = f32_to_u8(c.w);
            return 4;
        case ORC_FMT_BGRA8_UNORM: case ORC_FMT_BGRA8_UNORM_SRGB:
            out[0] = f32_to_u8(c.z); out[1] = f32_to_u8(c.y); out[2] = f32_to_u8(c.x); out[3] = f32_to_u8(c.w);
            return 4;
        default: return 0;
    }
}
/* ORC_EXT_BLEND (not in the reference; fragment.rs:480-485 ignores the blend state): the WebGPU blend equation */
static inline float srgb_eotf(float v) {
    if (!(v > 0.04045f)) return v / 12.92f;
    return powf((v + 0.055f) / 1.055f, 2.4f);
}
static Vec4 decode_color(const uint8_t* texel, uint32_t format, bool srgb) {
    float c[4] = {0, 0, 0, 0};
    const int bpp = bytes_per_texel(format);
    for (int k = 0; k < bpp && k < 4; k++) c[k] = (float)texel[k] / 255.0f;
    if (format == ORC_FMT_BGRA8_UNORM || format == ORC_FMT_BGRA8_UNORM_SRGB) { const float t = c[0]; c[0] = c[2]; c[2] = t; }
    if (srgb && (format == ORC_FMT_RGBA8_UNORM_SRGB || format == ORC_FMT_BGRA8_UNORM_SRGB)) { c[0] = srgb_eotf(c[0]); c[1] = srgb_eotf(c[1]); c[2] = srgb_eotf(c[2]); }
    return {c[0], c[1], c[2], c[3]};
}
static float blend_factor(uint32_t f, float s, float sa, float dv, float da, float c, bool alpha) {
    switch (f) {
        case 0: return 0.0f; case 1: return 1.0f;
        case 2: return s; case 3: return 1.0f - s;
        case 4: return sa; case 5: return 1.0f - sa;
        case 6: return dv; case 7: return 1.0f - dv;
        case 8: return da; case 9: return 1.0f - da;
        case 10: return alpha ? 1.0f : fminf(sa, 1.0f - da);
        case 11: return c;
        default: return 1.0f - c;
    }
}
static float blend_channel(uint32_t sf, uint32_t df, uint32_t op, float s, float sa, float dv, float da, float c, bool alpha) {
    if (op == 3) return fminf(s, dv);
    if (op == 4) return fmaxf(s, dv);
    const float a = s * blend_factor(sf, s, sa, dv, da, c, alpha);
    const float b = dv * blend_factor(df, s, sa, dv, da, c, alpha);
    return op == 0 ? a + b : op == 1 ? a - b : b - a;
}
static inline float clamp01(float v) { return fminf(fmaxf(v, 0.0f), 1.0f); }
static Vec4 blend_color(const uint32_t st[7], const float constant[4], Vec4 src, Vec4 dst) {
    src = {clamp01(src.x), clamp01(src.y), clamp01(src.z), clamp01(src.w)};
    Vec4 r;
    r.x = blend_channel(st[1], st[2], st[3], src.x, src.w, dst.x, dst.w, constant[0], false);
    r.y = blend_channel(st[1], st[2], st[3], src.y, src.w, dst.y, dst.w, constant[1], false);
    r.z = blend_channel(st[1], st[2], st[3], src.z, src.w, dst.z, dst.w, constant[2], false);
    r.w = blend_channel(st[4], st[5], st[6], src.w, src.w, dst.w, dst.w, constant[3], true);
    return r;
}
static inline uint64_t texel_offset(const orc_texture& t, uint32_t x, uint32_t y, int bpp) {
    /* TextureDataLayout::texel_byte_offset (texture.rs:283-285): x*bpp + y*W*bpp (+ z*W*H*bpp) */
    return (uint64_t)x * bpp + (uint64_t)y * t.width * bpp;
}

/* binding.rs:151-164 texel_coordinate */
static inline float rem_euclid(float x, float rhs) {   /* Rust f32::rem_euclid */
    const float r = fmodf(x, rhs);
    return r < 0.0f ? r + fabsf(rhs) : r;
}
static inline uint32_t texel_coordinate(float x, uint32_t mode, uint32_t size) {
    switch (mode) {
        case ORC_ADDR_CLAMP_TO_EDGE:
            /* f32::clamp(0,1): NaN propagates */
            if (x < 0.0f) x = 0.0f;
            if (x > 1.0f) x = 1.0f;
            break;
        case ORC_ADDR_REPEAT: x = rem_euclid(x, 1.0f); break;
        default: {
            const float r = rem_euclid(x, 2.0f);
            x = (r <= 1.0f) ? r : 2.0f - r;
        }
    }
    /* f32::round = half away from zero, as roundf; `as u32` saturates */
    return f32_as_u32(roundf(x * (float)(size - 1)));
}

const uint8_t* Resources::buffer(uint32_t g, uint32_t b, uint64_t min_size, int* err) const {
    if (g >= ORC_MAX_GROUPS || b >= ORC_MAX_BINDINGS || bindings->b[g][b].kind != ORC_BIND_BUFFER ||
        bindings->b[g][b].buffer.size < min_size) {
        *err = ORC_ERR_INVALID;
        return nullptr;
    }
    return bindings->b[g][b].buffer.data;
}
Vec4 Resources::sample(uint32_t tg, uint32_t tb, uint32_t sg, uint32_t sb, float u, float v, int* err) const {
    if (tg >= ORC_MAX_GROUPS || tb >= ORC_MAX_BINDINGS || sg >= ORC_MAX_GROUPS || sb >= ORC_MAX_BINDINGS ||
        bindings->b[tg][tb].kind != ORC_BIND_TEXTURE || bindings->b[sg][sb].kind != ORC_BIND_SAMPLER) {
        *err = ORC_ERR_INVALID;
        return {0, 0, 0, 0};
    }
    const orc_texture& img = bindings->b[tg][tb].texture;
    const orc_sampler& smp = bindings->b[sg][sb].sampler;
    if (img.format != ORC_FMT_RGBA8_UNORM && img.format != ORC_FMT_RGBA8_UNORM_SRGB) {
        *err = ORC_ERR_UNSUPPORTED;          /* texture.rs:186 `_ => todo!()` */
        return {0, 0, 0, 0};
    }
    const uint32_t tx = texel_coordinate(u, smp.address_u, img.width);
    const uint32_t ty = texel_coordinate(v, smp.address_v, img.height);
    if (tx >= img.width || ty >= img.height) { *err = ORC_ERR_OUT_OF_BOUNDS; return {0, 0, 0, 0}; }
    const uint8_t* p = img.data + texel_offset(img, tx, ty, 4);
    /* get_pixel (texture.rs:170-188): u8 as f32 / 255.0, no sRGB decode */
    return {(float)p[0] / 255.0f, (float)p[1] / 255.0f, (float)p[2] / 255.0f, (float)p[3] / 255.0f};
}

/* render_pass/mod.rs:434-452 */
static inline bool compare(uint32_t f, float value, float reference) {
    switch (f) {
        case ORC_CMP_NEVER: return false;
        case ORC_CMP_LESS: return value < reference;
        case ORC_CMP_EQUAL: return value == reference;
        case ORC_CMP_LESS_EQUAL: return value <= reference;
        case ORC_CMP_GREATER: return value > reference;
        case ORC_CMP_NOT_EQUAL: return value != reference;
        case ORC_CMP_GREATER_EQUAL: return value >= reference;
        default: return true;
    }
}

/* ------------------------------------------------------------------------- */
/* the draw: state.rs:238-593 + vertex.rs + fragment.rs                        */
/* ------------------------------------------------------------------------- */
struct VertexOut {       /* vertex.rs:204-210 VertexOutput */
    Vec4 clip;
    uint8_t inter[64];
};

struct DrawCtx {
    const orc_pass* pass;
    const orc_pipeline* pipe;
    const orc_draw* draw;
    const ShaderInfo* sh;
    Resources res;
    ToRaster target;
    orc_stats* stats;
    uint32_t* coverage;
    int err;
    int color_bpp[ORC_MAX_COLOR];
    const orc_raster_state* rs;
};

/* vertex.rs:286-316 VertexProcessingState::process + VertexInput::write_into (:128-156) */
static bool process_vertex(DrawCtx& cx, uint32_t instance_index, uint32_t vertex_index, VertexOut& out) {
    VsIn in;
    in.vertex_index = vertex_index;
    in.instance_index = instance_index;
    for (int i = 0; i < ORC_MAX_ATTRS; i++) { in.attr[i] = nullptr; in.attr_size[i] = 0; }
    for (uint32_t a = 0; a < cx.pipe->num_attrs; a++) {
        const orc_vertex_attr& at = cx.pipe->attrs[a];
        if (at.location >= ORC_MAX_ATTRS || at.buffer >= cx.pipe->num_vertex_buffers) { cx.err = ORC_ERR_INVALID; return false; }
        const orc_vertex_buffer_layout& l = cx.pipe->vb[at.buffer];
        const orc_buffer& buf = cx.draw->vertex_buffers[at.buffer];
        /* VertexBufferInput::read (vertex.rs:195-199) */
        const uint64_t stride_instance = l.step_mode == ORC_STEP_INSTANCE ? l.stride : 0;
        const uint64_t stride_vertex = l.step_mode == ORC_STEP_VERTEX ? l.stride : 0;
        const uint64_t start = (uint64_t)instance_index * stride_instance + (uint64_t)vertex_index * stride_vertex;
        if (buf.data == nullptr || start + at.offset + at.size > buf.size) { cx.err = ORC_ERR_OUT_OF_BOUNDS; return false; }
        in.attr[at.location] = buf.data + start + at.offset;
        in.attr_size[at.location] = at.size;
    }
    VsOut vo;
    std::memset(&vo, 0, sizeof(vo));
    int err = 0;
    cx.sh->vs(in, vo, cx.res, &err);
    if (err) { cx.err = err; return false; }
    out.clip = vo.position;
    std::memcpy(out.inter, vo.inter, sizeof(out.inter));
    return true;
}

/* fragment.rs:319-377 interpolate_user + util/interpolation.rs:149-163.
 * `coeff` has n entries; n == 1 selects vertex 0 without arithmetic (NoInterpolation). */
static bool build_varyings(DrawCtx& cx, int n, const float* coeff, const float* frag_w, const VertexOut* const* vin, uint8_t out[64]) {
    std::memset(out, 0, 64);
    for (int l = 0; l < cx.sh->num_varyings; l++) {
        const VaryingLayout& vl = cx.sh->varyings[l];
        if (vl.kind != VAR_F32 || vl.interp == INTERP_FLAT) {
            if (vl.interp != INTERP_FLAT) { cx.err = ORC_ERR_INVALID; return false; }   /* ints must be flat */
            std::memcpy(out + vl.offset, vin[0]->inter + vl.offset, 4 * vl.ncomp);       /* FlatSampling: vertex 0 */
            continue;
        }
        if (vl.interp == INTERP_PERSPECTIVE) {
            /* NOT IN THE REFERENCE (fragment.rs:343-345 is todo!()): WebGPU's perspective-correct interpolation,
             * sum(v_k B_k w_k) / sum(B_k w_k) with w_k = 1 / clip.w of the rasterised vertices (the fourth component of
             * to_raster's fragment, raster.rs:156) -- every product and sum a separate f32 operation, left to right */
            float bw[3] = {0.0f, 0.0f, 0.0f};
            for (int k = 0; k < n; k++) bw[k] = coeff[k] * frag_w[k];
            for (uint32_t c = 0; c < vl.ncomp; c++) {
                float p[3];
                for (int k = 0; k < n; k++) std::memcpy(&p[k], vin[k]->inter + vl.offset + 4 * c, 4);
                float value;
                if (n == 1) value = p[0];
                else {
                    float num = p[0] * bw[0], den = 1.0f * bw[0];
                    for (int k = 1; k < n; k++) { num = num + p[k] * bw[k]; den = den + 1.0f * bw[k]; }
                    value = num / den;
                }
                std::memcpy(out + vl.offset + 4 * c, &value, 4);
            }
            continue;
        }
        for (uint32_t c = 0; c < vl.ncomp; c++) {
            float p[3];
            for (int k = 0; k < n; k++) std::memcpy(&p[k], vin[k]->inter + vl.offset + 4 * c, 4);
            float accu;
            if (n == 1) accu = p[0];
            else {
                accu = p[0] * coeff[0];
                for (int k = 1; k < n; k++) accu = accu + p[k] * coeff[k];
            }
            std::memcpy(out + vl.offset + 4 * c, &accu, 4);
        }
    }
    return true;
}

/* FragmentOutput::depth_test (fragment.rs:427-453) */
static bool depth_test(DrawCtx& cx, uint32_t x, uint32_t y, float frag_depth) {
    if (cx.pass->has_depth && cx.pipe->has_depth_state) {
        const orc_texture& d = cx.pass->depth;
        float* slot = reinterpret_cast<float*>(d.data + texel_offset(d, x, y, 4));
        float stored;
        std::memcpy(&stored, slot, 4);
        if (compare(cx.pipe->depth_compare, frag_depth, stored)) {
            if (cx.pipe->depth_write) std::memcpy(slot, &frag_depth, 4);
            return true;
        }
        return false;
    }
    return true;
}

/* FragmentProcessingState::process (fragment.rs:127-214) */
static void process_fragment(DrawCtx& cx, int n, const VertexOut* const* unclipped, bool front_facing,
                             uint32_t primitive_index, uint32_t fx, uint32_t fy, Vec4 fragment, const float* coeff, const float* frag_w) {
    if (cx.err) return;
    for (uint32_t c = 0; c < cx.pass->num_color; c++)
        if (fx >= cx.pass->color[c].width || fy >= cx.pass->color[c].height) { cx.err = ORC_ERR_OUT_OF_BOUNDS; return; }
    if (cx.pass->has_depth && (fx >= cx.pass->depth.width || fy >= cx.pass->depth.height)) { cx.err = ORC_ERR_OUT_OF_BOUNDS; return; }
    if (cx.coverage) cx.coverage[(uint64_t)fy * cx.pass->color[0].width + fx] += 1;

    FsIn in;
    in.position = fragment;
    in.front_facing = front_facing;
    in.primitive_index = primitive_index;
    in.sample_index = 0;
    in.sample_mask = ~0u;
    if (!build_varyings(cx, n, coeff, frag_w, unclipped, in.inter)) return;

    float frag_depth = fragment.z;
    bool have_result = false, result = false;
    const int edt = cx.sh->early_depth_test;
    const bool early = edt != 0, late = edt != 1;
    if (early) {                                       /* :166-194 */
        const bool r = depth_test(cx, fx, fy, frag_depth);
        if (!r) return;
        if (!late) { have_result = true; result = r; }
    }
    FsOut out;
    std::memset(&out, 0, sizeof(out));
    int err = 0;
    cx.sh->fs(in, out, cx.res, &err);
    cx.stats->fragments_shaded += 1;
    if (err) { cx.err = err; return; }
    if (out.killed) return;                            /* :204-208 */
    /* FragmentOutput::read_from (:457-488) */
    if (out.has_frag_depth) frag_depth = out.frag_depth;
    bool wrote = false;
    for (int k = 0; k < out.num_color; k++) {
        if (!have_result) { result = depth_test(cx, fx, fy, frag_depth); have_result = true; }
        if (result) {
            const uint32_t loc = out.color_location[k];
            if (loc >= cx.pass->num_color) { cx.err = ORC_ERR_INVALID; return; }
            const orc_texture& t = cx.pass->color[loc];
            uint8_t texel[4];
            const bool srgb = (cx.rs->ext_features & ORC_EXT_SRGB_ENCODE) != 0;
            Vec4 value = out.color[k];
            if ((cx.rs->ext_features & ORC_EXT_BLEND) && cx.rs->blend[loc][0]) {
                const int dbpp = bytes_per_texel(t.format);
                if (dbpp == 0) { cx.err = ORC_ERR_UNSUPPORTED; return; }
                value = blend_color(cx.rs->blend[loc], cx.rs->blend_constant, value,
                                    decode_color(t.data + texel_offset(t, fx, fy, dbpp), t.format, srgb));
            }
            const int bpp = encode_color(value, t.format, texel, srgb);
            if (bpp == 0) { cx.err = ORC_ERR_UNSUPPORTED; return; }
            uint8_t* dst = t.data + texel_offset(t, fx, fy, bpp);
            if (cx.rs->ext_features & ORC_EXT_COLOR_WRITE_MASK) {             /* not in the reference: channels outside the mask keep their value */
                const uint32_t m = cx.rs->color_write_mask[loc];
                const bool bgra = t.format == ORC_FMT_BGRA8_UNORM || t.format == ORC_FMT_BGRA8_UNORM_SRGB;
                for (int ch = 0; ch < bpp; ch++) {
                    const int bit = (bgra && ch == 0) ? 2 : (bgra && ch == 2) ? 0 : ch;    /* stored byte -> R G B A */
                    if (!((m >> bit) & 1u)) texel[ch] = dst[ch];
                }
            }
            std::memcpy(dst, texel, bpp);                                     /* put_pixel texture.rs:213-218 */
            wrote = true;
        }
    }
    if (wrote) cx.stats->fragments_written += 1;
}

/* TriFace::new + front_face (primitive.rs:169-197); false in *ok if a w is 0 (reference panics) */
static bool tri_is_ccw(const Vec4 c[3], bool* ok) {
    float n[3][3];
    for (int i = 0; i < 3; i++) {
        if (c[i].w == 0.0f) { *ok = false; return false; }
        n[i][0] = c[i].x / c[i].w; n[i][1] = c[i].y / c[i].w; n[i][2] = c[i].z / c[i].w;
    }
    *ok = true;
    const float abx = n[1][0] - n[0][0], aby = n[1][1] - n[0][1];
    const float acx = n[2][0] - n[0][0], acy = n[2][1] - n[0][1];
    const float area_z = abx * acy - aby * acx;       /* z of ab.cross(ac) */
    return area_z > 0.0f;
}

static inline float shoelace(float ax, float ay, float bx, float by, float cx, float cy) {   /* raster.rs:293-297 */
    return (by - ay) * (bx + ax) + (cy - by) * (cx + bx) + (ay - cy) * (ax + cx);
}

/* TriRasterizer::rasterize (raster.rs:280-363) feeding FragmentProcessingState::process */
static void rasterize_tri(DrawCtx& cx, const ClipTri& tri, const VertexOut* const* unclipped, bool front_facing,
                          uint32_t primitive_index) {
    uint32_t fb[3][2];
    Vec4 frag[3];
    for (int i = 0; i < 3; i++)
        if (!cx.target.to_raster(tri.v[i].position, fb[i], frag[i])) { cx.err = ORC_ERR_W_ZERO; return; }
    const float total_area = shoelace(frag[0].x, frag[0].y, frag[1].x, frag[1].y, frag[2].x, frag[2].y);
    if (total_area == 0.0f) return;                    /* :306-310 */
    scanlines(fb, [&](const Row& row) {
        if (cx.err) return;
        if (row.y < 0 || row.x1 < 0 || row.y > 0xFFFFFFFFll || row.x2 > 0xFFFFFFFFll) { cx.err = ORC_ERR_INVALID; return; }
        for (int64_t xi = row.x1; xi <= row.x2; xi++) {
            const uint32_t x = (uint32_t)xi, y = (uint32_t)row.y;
            if (!cx.target.in_scissor(x, y)) continue;                                 /* :326-328 */
            const float px = (float)x, py = (float)y;
            const float b0 = shoelace(px, py, frag[1].x, frag[1].y, frag[2].x, frag[2].y) / total_area;
            const float b1 = shoelace(frag[0].x, frag[0].y, px, py, frag[2].x, frag[2].y) / total_area;
            const float b2 = shoelace(frag[0].x, frag[0].y, frag[1].x, frag[1].y, px, py) / total_area;
            /* compose with the clip barycentrics (:338-343): B' = I0*b0 + I1*b1 + I2*b2 */
            Bary3 B = b3_scale(tri.v[0].barycentric, b0);
            B = b3_add(B, b3_scale(tri.v[1].barycentric, b1));
            B = b3_add(B, b3_scale(tri.v[2].barycentric, b2));
            /* :345-346 -- the reference interpolates the CLIPPED vertices' fragment with
             * coefficients that refer to the UNCLIPPED vertices; mirrored */
            Vec4 f = v4_scale(frag[0], B.c[0]);
            f = v4_add(f, v4_scale(frag[1], B.c[1]));
            f = v4_add(f, v4_scale(frag[2], B.c[2]));
            const float fw[3] = {frag[0].w, frag[1].w, frag[2].w};
            process_fragment(cx, 3, unclipped, front_facing, primitive_index, x, y, f, B.c, fw);
        }
    });
}

/* LineRasterizer::rasterize (raster.rs:233-264) */
static void rasterize_line(DrawCtx& cx, const Vec4 clipped[2], const float alphas[2], const VertexOut* const* unclipped,
                           uint32_t primitive_index) {
    uint32_t s[2], e[2];
    Vec4 fs, fe;
    if (!cx.target.to_raster(clipped[0], s, fs) || !cx.target.to_raster(clipped[1], e, fe)) { cx.err = ORC_ERR_W_ZERO; return; }
    Bresenham br(s, e);
    uint32_t pt[2];
    float t;
    while (br.next(pt, &t)) {
        if (cx.err) return;
        if (!cx.target.in_scissor(pt[0], pt[1])) continue;
        /* Lerp(t).interpolate([Lerp(a0), Lerp(a1)]) = Lerp(a0*(1-t) + a1*t) */
        const float tt = alphas[0] * (1.0f - t) + alphas[1] * t;
        const float c[2] = {1.0f - tt, tt};
        const Vec4 f = v4_add(v4_scale(fs, c[0]), v4_scale(fe, c[1]));
        const float fw[2] = {fs.w, fe.w};
        process_fragment(cx, 2, unclipped, true, primitive_index, pt[0], pt[1], f, c, fw);
    }
}

/* PointRasterizer::rasterize (raster.rs:197-217) */
static void rasterize_point(DrawCtx& cx, const VertexOut* const* unclipped, uint32_t primitive_index) {
    uint32_t p[2];
    Vec4 f;
    if (!cx.target.to_raster(unclipped[0]->clip, p, f)) { cx.err = ORC_ERR_W_ZERO; return; }
    if (!cx.target.in_scissor(p[0], p[1])) return;
    const float c[1] = {1.0f};
    const float fw[1] = {f.w};
    process_fragment(cx, 1, unclipped, true, primitive_index, p[0], p[1], f, c, fw);
}

struct Item { bool separator; VertexOut v; };

static int draw_execute(DrawCtx& cx) {
    const orc_pipeline& pipe = *cx.pipe;
    const orc_draw& dr = *cx.draw;
    const bool is_strip = pipe.topology == ORC_TOPO_LINE_STRIP || pipe.topology == ORC_TOPO_TRIANGLE_STRIP;
    bool separated = false;
    if (dr.indexed) {
        if (pipe.strip_index_format != ORC_INDEX_NONE) {          /* state.rs:296-305, :357-359 */
            if (!is_strip) return ORC_ERR_INVALID;
            if (pipe.strip_index_format != dr.index_format) return ORC_ERR_INVALID;
            separated = true;
        }
        if (dr.index_format != ORC_INDEX_U16 && dr.index_format != ORC_INDEX_U32) return ORC_ERR_INVALID;
    }
    const int prim_size = pipe.topology == ORC_TOPO_POINT_LIST ? 1
                        : (pipe.topology == ORC_TOPO_LINE_LIST || pipe.topology == ORC_TOPO_LINE_STRIP) ? 2 : 3;
    std::vector<Item> items;
    std::deque<ClipTri> queue;
    items.reserve(dr.count);

    for (uint32_t ii = 0; ii < dr.instance_count; ii++) {          /* state.rs:519 */
        const uint32_t instance_index = dr.first_instance + ii;
        items.clear();
        /* index resolution + vertex processing, once per index, no reuse (state.rs:521-535) */
        for (uint32_t k = 0; k < dr.count; k++) {
            const uint32_t index = dr.first + k;
            Item it;
            it.separator = false;
            uint32_t vertex = index;                                /* DirectIndices index.rs:27-33 */
            if (dr.indexed) {
                int sep = 0;
                const int e = orc_resolve_index(dr.index_buffer.data, dr.index_buffer.size, dr.index_format,
                                                dr.base_vertex, index, separated ? 1 : 0, &vertex, &sep);
                if (e) return e;
                it.separator = sep != 0;
            }
            cx.stats->vertices_processed += 1;
            if (!it.separator) {
                if (!process_vertex(cx, instance_index, vertex, it.v)) return cx.err;
            }
            items.push_back(it);
        }

        /* primitive assembly (primitive.rs:216-487) -> lists of item indices */
        std::vector<uint32_t> prims;     /* prim_size indices into `items` per primitive */
        const uint32_t n = (uint32_t)items.size();
        if (!is_strip) {
            for (uint32_t k = 0; k + prim_size <= n; k += prim_size)      /* array_chunks_: incomplete tail dropped */
                for (int j = 0; j < prim_size; j++) prims.push_back(k + j);
        } else if (prim_size == 2) {
            /* Strip<2> (:281-331): windows, cleared at separators */
            int have = -1;
            for (uint32_t k = 0; k < n; k++) {
                if (items[k].separator) { have = -1; continue; }
                if (have >= 0) { prims.push_back((uint32_t)have); prims.push_back(k); }
                have = (int)k;
            }
        } else {
            /* Strip<3> + TriStripGenerator (:333-487) */
            uint32_t buf[2]; int filled = 0; bool odd = false;
            for (uint32_t k = 0; k < n; k++) {
                if (items[k].separator) { filled = 0; odd = false; continue; }
                if (filled < 2) { buf[filled++] = k; continue; }
                prims.push_back(buf[0]); prims.push_back(buf[1]); prims.push_back(k);
                buf[odd ? 1 : 0] = k;
                odd = !odd;
            }
        }

        if (!pipe.has_fragment) continue;                             /* state.rs:583-588 */

        const uint32_t nprims = (uint32_t)(prims.size() / prim_size);
        for (uint32_t primitive_index = 0; primitive_index < nprims; primitive_index++) {   /* :541 enumerate */
            cx.stats->primitives_assembled += 1;
            const VertexOut* v[3] = {nullptr, nullptr, nullptr};
            for (int j = 0; j < prim_size; j++) v[j] = &items[prims[primitive_index * prim_size + j]].v;

            if (prim_size == 3) {
                const Vec4 pos[3] = {v[0]->clip, v[1]->clip, v[2]->clip};
                bool ok;
                const bool ccw = tri_is_ccw(pos, &ok);                 /* computed at assembly, before culling */
                if (!ok) return ORC_ERR_W_ZERO;
                const bool front_facing = (ccw ? ORC_FRONT_CCW : ORC_FRONT_CW) == (int)pipe.front_face;   /* :546-547 */
                const bool cull = pipe.cull_mode == ORC_CULL_FRONT ? front_facing
                                : pipe.cull_mode == ORC_CULL_BACK ? !front_facing : false;
                if (cull) { cx.stats->primitives_culled += 1; continue; }
                clip_tri(pos, queue);
                for (const ClipTri& t : queue) {
                    rasterize_tri(cx, t, v, front_facing, primitive_index);
                    if (cx.err) return cx.err;
                    cx.stats->primitives_drawn += 1;
                }
            } else if (prim_size == 2) {
                Vec4 pts[2] = {v[0]->clip, v[1]->clip};
                float alphas[2];
                if (clip_line(pts, alphas)) {
                    rasterize_line(cx, pts, alphas, v, primitive_index);
                    if (cx.err) return cx.err;
                    cx.stats->primitives_drawn += 1;
                }
            } else {
                rasterize_point(cx, v, primitive_index);               /* NoClipper: points are never clipped */
                if (cx.err) return cx.err;
                cx.stats->primitives_drawn += 1;
            }
        }
    }
    return ORC_OK;
}

}  // namespace orc

/* ========================================================================= */
/* C API                                                                       */
/* ========================================================================= */
using namespace orc;

extern "C" {

int orc_pass_load(const orc_pass* pass) {
    /* AcquiredColorAttachment::load -> clear_color (fragment.rs:537-549, texture.rs:199-206) */
    for (uint32_t c = 0; c < pass->num_color; c++) {
        if (!pass->color_clear[c]) continue;
        const orc_texture& t = pass->color[c];
        /* wgpu_color_to_vec4: f64 -> f32 cast (texture.rs:505-507) */
        Vec4 col = {(float)pass->clear_color[c][0], (float)pass->clear_color[c][1],
                    (float)pass->clear_color[c][2], (float)pass->clear_color[c][3]};
        uint8_t texel[4];
        const int bpp = encode_color(col, t.format, texel, (pass->ext_features & ORC_EXT_SRGB_ENCODE) != 0);
        if (bpp == 0) return ORC_ERR_UNSUPPORTED;
        const uint64_t n = (uint64_t)t.width * t.height;
        for (uint64_t i = 0; i < n; i++) std::memcpy(t.data + i * bpp, texel, bpp);
    }
    /* AcquiredDepthStencilAttachment::load -> clear_depth: fill(f32) (fragment.rs:603-621, texture.rs:208-211) */
    if (pass->has_depth && pass->depth_clear) {
        if (pass->depth.format != ORC_FMT_DEPTH32_FLOAT) return ORC_ERR_UNSUPPORTED;
        float* d = reinterpret_cast<float*>(pass->depth.data);
        const uint64_t n = (uint64_t)pass->depth.width * pass->depth.height;
        for (uint64_t i = 0; i < n; i++) d[i] = pass->clear_depth;
    }
    return ORC_OK;
}

void orc_default_raster_state(uint32_t width, uint32_t height, orc_raster_state* out) {
    out->vp_x = 0.0f; out->vp_y = 0.0f;
    out->vp_w = (float)width; out->vp_h = (float)height;
    out->vp_min_depth = 0.0f; out->vp_max_depth = 1.0f;
    out->sc_x = 0; out->sc_y = 0; out->sc_w = width; out->sc_h = height;
    out->ext_features = 0;
    for (int c = 0; c < ORC_MAX_COLOR; c++) out->color_write_mask[c] = 15u;
    std::memset(out->blend, 0, sizeof(out->blend));
    for (int k = 0; k < 4; k++) out->blend_constant[k] = 0.0f;
}

int orc_draw_execute(const orc_pass* pass, const orc_pipeline* pipe, const orc_raster_state* rs,
                     const orc_bindings* bindings, const orc_draw* draw, orc_stats* stats, uint32_t* coverage) {
    const ShaderInfo* sh = shader_info(pipe->shader);
    if (!sh) return ORC_ERR_INVALID;
    /* State::new asserts all attachments have the same size (state.rs:84-96) */
    uint32_t w = 0, h = 0; bool have = false;
    for (uint32_t c = 0; c < pass->num_color; c++) {
        if (have && (pass->color[c].width != w || pass->color[c].height != h)) return ORC_ERR_INVALID;
        w = pass->color[c].width; h = pass->color[c].height; have = true;
    }
    if (pass->has_depth) {
        if (have && (pass->depth.width != w || pass->depth.height != h)) return ORC_ERR_INVALID;
        if (pass->depth.format != ORC_FMT_DEPTH32_FLOAT) return ORC_ERR_UNSUPPORTED;
    }
    if (coverage && pass->num_color == 0) return ORC_ERR_INVALID;
    orc_stats local;
    std::memset(&local, 0, sizeof(local));
    DrawCtx cx{pass, pipe, draw, sh, Resources{bindings}, ToRaster(*rs), stats ? stats : &local, coverage, 0, {0, 0, 0, 0}, rs};
    const int e = draw_execute(cx);
    return e ? e : cx.err;
}

void orc_bubblesort3_i64(int64_t v[3]) {
    bubblesort3(v, [](int64_t a, int64_t b) { return a > b; });
}

int orc_scanlines(const uint32_t tri[6], uint32_t* rows, int max_rows) {
    uint32_t fb[3][2] = {{tri[0], tri[1]}, {tri[2], tri[3]}, {tri[4], tri[5]}};
    int n = 0;
    scanlines(fb, [&](const Row& r) {
        if (n < max_rows) { rows[3 * n] = (uint32_t)r.y; rows[3 * n + 1] = (uint32_t)r.x1; rows[3 * n + 2] = (uint32_t)r.x2; }
        n++;
    });
    return n;
}

int orc_clip_triangle(const float in[12], float* out_pos, float* out_bary, int max_tris) {
    Vec4 pos[3];
    for (int i = 0; i < 3; i++) pos[i] = {in[4 * i], in[4 * i + 1], in[4 * i + 2], in[4 * i + 3]};
    std::deque<ClipTri> q;
    clip_tri(pos, q);
    int n = 0;
    for (const ClipTri& t : q) {
        if (n < max_tris)
            for (int i = 0; i < 3; i++) {
                float* p = out_pos + (n * 3 + i) * 4;
                p[0] = t.v[i].position.x; p[1] = t.v[i].position.y; p[2] = t.v[i].position.z; p[3] = t.v[i].position.w;
                for (int k = 0; k < 3; k++) out_bary[(n * 3 + i) * 3 + k] = t.v[i].barycentric.c[k];
            }
        n++;
    }
    return n;
}

int orc_clip_line(const float in[8], float out_pos[8], float out_alpha[2]) {
    Vec4 pts[2] = {{in[0], in[1], in[2], in[3]}, {in[4], in[5], in[6], in[7]}};
    if (!clip_line(pts, out_alpha)) return 0;
    for (int i = 0; i < 2; i++) { out_pos[4 * i] = pts[i].x; out_pos[4 * i + 1] = pts[i].y; out_pos[4 * i + 2] = pts[i].z; out_pos[4 * i + 3] = pts[i].w; }
    return 1;
}

int orc_bresenham(const uint32_t start[2], const uint32_t end[2], uint32_t* pts, float* ts, int max_pts) {
    Bresenham b(start, end);
    int n = 0;
    uint32_t p[2]; float t;
    while (b.next(p, &t)) {
        if (n < max_pts) { pts[2 * n] = p[0]; pts[2 * n + 1] = p[1]; ts[n] = t; }
        n++;
    }
    return n;
}

int orc_tri_strip(const uint32_t* items, int n, int separated, uint32_t* out_tris, int max_tris) {
    uint32_t buf[2]; int filled = 0; bool odd = false; int m = 0;
    for (int k = 0; k < n; k++) {
        if (separated && items[k] == 0xFFFFFFFFu) { filled = 0; odd = false; continue; }
        if (filled < 2) { buf[filled++] = items[k]; continue; }
        if (m < max_tris) { out_tris[3 * m] = buf[0]; out_tris[3 * m + 1] = buf[1]; out_tris[3 * m + 2] = items[k]; }
        m++;
        buf[odd ? 1 : 0] = items[k];
        odd = !odd;
    }
    return m;
}

int orc_front_face_ccw(const float clip[12]) {
    Vec4 c[3];
    for (int i = 0; i < 3; i++) c[i] = {clip[4 * i], clip[4 * i + 1], clip[4 * i + 2], clip[4 * i + 3]};
    bool ok;
    const bool ccw = tri_is_ccw(c, &ok);
    return ok ? (ccw ? 1 : 0) : -1;
}

/* IndirectIndices::lookup + to_vertex (index.rs:45-88): u32::from(index).strict_add_signed(base_vertex) */
int orc_resolve_index(const uint8_t* index_data, uint64_t size, uint32_t format, int32_t base_vertex,
                      uint32_t i, int separated, uint32_t* out_vertex, int* is_separator) {
    uint32_t raw;
    *is_separator = 0;
    if (format == ORC_INDEX_U16) {
        if (((uint64_t)i + 1) * 2 > size) return ORC_ERR_OUT_OF_BOUNDS;
        uint16_t v; std::memcpy(&v, index_data + (uint64_t)i * 2, 2);
        if (separated && v == 0xFFFFu) { *is_separator = 1; return ORC_OK; }
        raw = v;
    } else if (format == ORC_INDEX_U32) {
        if (((uint64_t)i + 1) * 4 > size) return ORC_ERR_OUT_OF_BOUNDS;
        uint32_t v; std::memcpy(&v, index_data + (uint64_t)i * 4, 4);
        if (separated && v == 0xFFFFFFFFu) { *is_separator = 1; return ORC_OK; }
        raw = v;
    } else return ORC_ERR_INVALID;
    const int64_t r = (int64_t)raw + (int64_t)base_vertex;
    if (r < 0 || r > 0xFFFFFFFFll) return ORC_ERR_OUT_OF_BOUNDS;      /* strict_add_signed panics on overflow */
    *out_vertex = (uint32_t)r;
    return ORC_OK;
}

int orc_to_raster(const orc_raster_state* rs, const float clip[4], uint32_t fb[2], float frag[4]) {
    ToRaster t(*rs);
    Vec4 f;
    if (!t.to_raster({clip[0], clip[1], clip[2], clip[3]}, fb, f)) return ORC_ERR_W_ZERO;
    frag[0] = f.x; frag[1] = f.y; frag[2] = f.z; frag[3] = f.w;
    return ORC_OK;
}

uint32_t orc_texel_coordinate(float x, uint32_t address_mode, uint32_t size) { return texel_coordinate(x, address_mode, size); }
uint8_t orc_f32_to_u8(float v) { return f32_to_u8(v); }

uint64_t orc_texture_byte_size(uint32_t format, uint32_t w, uint32_t h, uint32_t layers) {
    /* TextureDataLayout::from_size (texture.rs:257-273): stride = (bpp, W*bpp, H*W*bpp, L*H*W*bpp) */
    return (uint64_t)bytes_per_texel(format) * w * h * layers;
}
uint64_t orc_texel_byte_offset(uint32_t format, uint32_t w, uint32_t h, uint32_t x, uint32_t y, uint32_t z) {
    const uint64_t bpp = bytes_per_texel(format);
    return x * bpp + (uint64_t)y * w * bpp + (uint64_t)z * w * h * bpp;
}

int orc_run_vertex_shader(uint32_t shader, uint32_t vertex_index, uint32_t instance_index,
                          const orc_bindings* bindings, float out_position[4], float out_varyings[16]) {
    const ShaderInfo* sh = shader_info(shader);
    if (!sh) return ORC_ERR_INVALID;
    VsIn in;
    in.vertex_index = vertex_index; in.instance_index = instance_index;
    for (int i = 0; i < ORC_MAX_ATTRS; i++) { in.attr[i] = nullptr; in.attr_size[i] = 0; }
    VsOut out;
    std::memset(&out, 0, sizeof(out));
    orc_bindings empty;
    std::memset(&empty, 0, sizeof(empty));
    Resources res{bindings ? bindings : &empty};
    int err = 0;
    sh->vs(in, out, res, &err);
    if (err) return err;
    out_position[0] = out.position.x; out_position[1] = out.position.y; out_position[2] = out.position.z; out_position[3] = out.position.w;
    std::memcpy(out_varyings, out.inter, 64);
    return ORC_OK;
}

}  // extern "C"
