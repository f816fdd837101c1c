/* oracle_internal.h -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 * Shader ABI between oracle.cpp (the path) and oracle_shaders.cpp (hand-written
 * restatements of the WGSL shaders in the op order the reference JIT emits). */
#ifndef WGPU_CPU_ORACLE_INTERNAL_H
#define WGPU_CPU_ORACLE_INTERNAL_H

#include "oracle.h"
#include <cstring>

namespace orc {

struct Vec4 { float x, y, z, w; };

/* one inter-stage location (naga-cranelift/src/bindings.rs:307-346: locations are
 * packed in visit order, each aligned to its naga alignment) */
enum { VAR_F32 = 0, VAR_U32 = 1, VAR_I32 = 2 };
enum { INTERP_FLAT = 0, INTERP_LINEAR = 1, INTERP_PERSPECTIVE = 2 };
struct VaryingLayout {
    uint32_t location;
    uint32_t offset;      /* byte offset in the inter-stage buffer */
    uint32_t ncomp;       /* 1..4 */
    uint32_t kind;        /* VAR_* */
    uint32_t interp;      /* INTERP_* */
};

struct VsIn {
    uint32_t vertex_index, instance_index;
    const uint8_t* attr[ORC_MAX_ATTRS];   /* by @location; null if not supplied */
    uint32_t attr_size[ORC_MAX_ATTRS];
};
struct VsOut {
    Vec4 position;
    uint8_t inter[64];
};
struct FsIn {
    Vec4 position;
    bool front_facing;
    uint32_t primitive_index, sample_index, sample_mask;
    uint8_t inter[64];
};
/* fragment.rs:457-488: outputs are visited in struct order; FragDepth replaces
 * frag_depth; the late depth test runs at the FIRST @location output. */
struct FsOut {
    bool killed;
    bool has_frag_depth; /* and it precedes the first location in struct order */
    float frag_depth;
    int num_color;
    Vec4 color[ORC_MAX_COLOR];
    uint32_t color_location[ORC_MAX_COLOR];
};

struct Resources {
    const orc_bindings* bindings;
    /* runtime.rs:523-561 buffer_resource */
    const uint8_t* buffer(uint32_t g, uint32_t b, uint64_t min_size, int* err) const;
    /* binding.rs:93-149 image_sample (nearest, mip 0, 2-D) */
    Vec4 sample(uint32_t tg, uint32_t tb, uint32_t sg, uint32_t sb, float u, float v, int* err) const;
};

typedef void (*VsFn)(const VsIn&, VsOut&, const Resources&, int* err);
typedef void (*FsFn)(const FsIn&, FsOut&, const Resources&, int* err);

struct ShaderInfo {
    VsFn vs;
    FsFn fs;
    int num_varyings;
    VaryingLayout varyings[8];
    /* naga EarlyDepthTest: 0 none (late only), 1 force (early only), 2 allow (both) */
    int early_depth_test;
};

const ShaderInfo* shader_info(uint32_t shader);

}  // namespace orc
#endif
