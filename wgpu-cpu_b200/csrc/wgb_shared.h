// wgb_shared.h -- plain-old-data shared by the host runtime (wgb_api.cpp) and the device
// code compiled per pipeline with NVRTC (wgb_prelude.cuh / wgb_raster.cuh).
#pragma once

#ifdef __CUDACC_RTC__
typedef unsigned char wgb_u8;
typedef unsigned short wgb_u16;
typedef unsigned int wgb_u32;
typedef int wgb_i32;
typedef unsigned long long wgb_u64;
#else
#include <stdint.h>
typedef uint8_t wgb_u8;
typedef uint16_t wgb_u16;
typedef uint32_t wgb_u32;
typedef int32_t wgb_i32;
typedef uint64_t wgb_u64;
#endif

#define WGB_MAX_GROUPS 4
#define WGB_MAX_BINDINGS 4
#define WGB_MAX_VERTEX_BUFFERS 8
#define WGB_MAX_COLOR 4

// screen tile processed by one CTA of the tile kernel
#ifndef WGB_TILE_W
#define WGB_TILE_W 32
#endif
#ifndef WGB_TILE_H
#define WGB_TILE_H 32
#endif
// a primitive whose tile bounding box holds more than this many tiles goes to the
// "big" list that every tile scans instead of into per-tile bins; this bounds the bin
// storage at WGB_SMALL_MAX_TILES entries per primitive
#define WGB_SMALL_MAX_TILES 4
// upper bound on sub-triangles the six-plane clipper can emit for one triangle
// (wgpu-cpu/src/render_pass/clipper.rs:737-739 "2**6")
#define WGB_MAX_CLIP_TRIS 64

// resource kinds in a bind slot
#define WGB_RES_NONE 0
#define WGB_RES_BUFFER 1
#define WGB_RES_TEXTURE 2
#define WGB_RES_SAMPLER 3

struct WgbResource {            // 32 bytes
    wgb_u64 ptr;                // buffer base (already offset) / texture linear base
    wgb_u64 tex;                // CUtexObject (bindless) for textures
    wgb_u32 a;                  // buffer: size in bytes; texture: width;  sampler: address_mode_u
    wgb_u32 b;                  //                         texture: height; sampler: address_mode_v
    wgb_u32 c;                  //                         texture: format
    wgb_u32 kind;
};

struct WgbAttachment {
    wgb_u64 ptr;                // linear, row-major, no padding (texture.rs:250-302)
    wgb_u32 format;
    wgb_u32 load_clear;         // 1: LoadOp::Clear still pending for this pass (first draw), 0: load stored texels
    wgb_u32 clear_texel;        // encoded clear colour (colour) / f32 bits (depth)
    wgb_u32 bytes_per_texel;
    wgb_u32 write_mask;         // opt-in colour write mask as a byte mask over the texel (0xFFFFFFFF = write everything)
    wgb_u32 srgb_encode;        // opt-in: apply the sRGB transfer function to r, g, b before the 8-bit encode
};

struct WgbVertexBuffer {
    wgb_u64 ptr;
    wgb_u64 size;
};

// counters + status written by the geometry kernels, read by the tile kernel and the host
struct WgbCounters {             // 64 bytes
    wgb_u32 num_slow;           // primitives that need the clipper
    wgb_u32 num_clip_records;   // sub-triangle records reserved by the clip kernel
    wgb_u32 num_big;            // entries in the big list
    wgb_u32 num_small_pairs;    // total (entry, tile) pairs in the bins (after the scan)
    wgb_u32 status;             // WGB_STATUS_* bits
    wgb_u32 max_tile_pairs;     // largest per-tile pair count the tile kernel saw (sizes bin_cap for the next draw)
    wgb_u64 fragments;          // rasterised fragments (what the reference runs its fragment stage on)
    wgb_u64 shaded;             // fragment-shader invocations for surviving fragments
    wgb_u32 hiz_culled;         // bin entries the tile kernel dropped by the hierarchical depth test
    wgb_u32 pad[5];
};
#define WGB_STATUS_CLIP_OVERFLOW 1u
#define WGB_STATUS_INDEX_OOB 2u
#define WGB_STATUS_VERTEX_OOB 4u
#define WGB_STATUS_W_ZERO 8u
#define WGB_STATUS_BIG_OVERFLOW 16u
#define WGB_STATUS_BIN_OVERFLOW 32u    // direct binning: a tile received more than bin_cap entries

// post-transform vertex flags: ONE BYTE per vertex (the geometry stage gathers three of them per primitive -- of every
// primitive, on every rank of a sort-first partition -- so the table is kept small enough to stay in L1 / L2: 5 MB at C3)
#define WGB_VFLAG_INSIDE 1u              // inside all six clip planes
#define WGB_VFLAG_W_ZERO 2u              // clip.w == 0 (the reference panics when such a vertex is used)
#define WGB_VFLAG_W_POS 4u               // clip.w > 0: clipping keeps the primitive inside the row range of its vertices
#define WGB_VFLAG_ROW_OK 8u              // the framebuffer row trunc(vp.y) is below 65535; the four bits below are valid
#define WGB_VFLAG_ABOVE 16u              // row < first row of the draw's window (scissor, framebuffer, this rank's band)
#define WGB_VFLAG_BELOW 32u              // row >= one past the window's last row
#define WGB_VFLAG_ABOVE1 64u             // row + 1 < first row   (one row of slack: clipped primitives)
#define WGB_VFLAG_BELOW1 128u            // row >= one past the last row + 1

// one record per primitive emitted by the clipper (slow path only)
struct WgbClipRecord {          // 100 bytes
    float frag[3][4];           // to_raster'd CLIPPED vertices: (vp.x, vp.y, ndc.z, 1/w)
    float bary[3][3];           // triangles: clip barycentrics w.r.t. the UNCLIPPED vertices; lines: [0][0..1] = alphas
    wgb_u32 prim;               // primitive sequence number within the batch
    wgb_u32 sub;                // index in the clipper's output queue
    wgb_u32 front_facing;
    wgb_u32 box;                // tile box, same encoding as prim_box
};

struct WgbBigEntry {            // 16 bytes
    wgb_u16 tx0, ty0, tx1, ty1; // inclusive tile bounding box
    wgb_u32 entry;              // bin entry encoding (see below)
    wgb_u32 pad;
};
// bin entry encoding: bit 31 = 1 -> clip record index in the low bits, 0 -> primitive sequence number
#define WGB_ENTRY_CLIP 0x80000000u

// a CUtensorMap (128 bytes, 64-byte aligned), encoded by the host with cuTensorMapEncodeTiled: the attachment as a
// 2-D tensor of 4-byte texels with a 32 x 32 box, for one-instruction tile loads / stores (cp.async.bulk.tensor.2d)
struct alignas(64) WgbTensorMap { wgb_u64 opaque[16]; };

// parameters of one draw batch; passed by value as a __grid_constant__ kernel argument
struct WgbDraw {
    // framebuffer
    wgb_u32 fb_width, fb_height;
    wgb_u32 tiles_x, tiles_y;            // whole framebuffer in tiles
    wgb_u32 band_ty0, band_ty1;          // this rank's tile-row band [ty0, ty1) (sort-first partition)
    // raster state (state.rs:604-628, raster.rs:129-143)
    float vp_tx, vp_ty, vp_sx, vp_sy;    // ToRaster translation / scaling
    float depth_min, depth_scale;        // opt-in viewport depth range: z = depth_min + ndc.z * depth_scale
    wgb_u32 depth_range;                 // 0: the reference's behaviour (ndc.z as it is, raster.rs:141-158)
    wgb_u32 pad3;
    float blend_constant[4];             // set_blend_constant (used by WGB_FEATURE_BLEND pipelines only)
    wgb_u32 sc_x0, sc_y0, sc_x1, sc_y1;  // scissor_bb
    // draw call (state.rs:225-236)
    wgb_u32 indexed;                     // 0 direct, 1 u16, 2 u32
    wgb_u32 first;                       // first index / first vertex
    wgb_u32 count;                       // indices / vertices per instance
    wgb_i32 base_vertex;
    wgb_u32 first_instance, instance_count;
    wgb_u32 prims_per_instance;          // assembled primitives per instance
    wgb_u32 prim_base;                   // first primitive sequence number of this batch
    wgb_u32 num_prims;                   // primitives in this batch
    wgb_u32 stats;                       // 1: maintain fragment counters + coverage buffer
    wgb_u64 index_ptr;
    wgb_u64 index_size;
    wgb_u64 strip_map;                   // separated strips: u32 x 3 vertex slots per primitive, else 0
    WgbVertexBuffer vb[WGB_MAX_VERTEX_BUFFERS];
    WgbResource res[WGB_MAX_GROUPS][WGB_MAX_BINDINGS];
    // attachments
    wgb_u32 num_color;
    wgb_u32 has_depth;
    WgbAttachment color[WGB_MAX_COLOR];
    WgbAttachment depth;
    // work buffers
    wgb_u64 counters;                    // WgbCounters*
    wgb_u64 prim_box;                    // u32 per primitive: kind | packed tile box / clip record base
    wgb_u64 setup_cache;                 // float4 x 3 per primitive: to_raster'd vertices (vp.x, vp.y, ndc.z, 1/w) of unclipped primitives
    // post-transform vertex cache (indexed draws): the vertex stage runs once per (instance, vertex) instead of once per index
    wgb_u64 vcache_raster;               // float4 per vertex: (vp.x, vp.y, ndc.z, 1/w)
    wgb_u64 vcache_ndc;                  // float2 per vertex: ndc.xy (winding)
    wgb_u64 vcache_flags;                // u8 per vertex: WGB_VFLAG_*
    wgb_u32 vcache_count;                // vertices per instance in the cache (0 = cache not in use)
    wgb_u32 pad2;
    wgb_u64 slow_list;                   // u32 per slow primitive
    wgb_u64 clip_records;                // WgbClipRecord*
    wgb_u32 clip_capacity;
    wgb_u32 pad1;
    wgb_u64 big_list;                    // WgbBigEntry*
    wgb_u32 big_capacity;
    wgb_u32 bin_cap;                     // direct binning: entries reserved per tile, bins[tile * bin_cap + slot]; 0 = count / scan / fill
    wgb_u64 tile_count;                  // u32 per tile in the band
    wgb_u64 tile_offset;                 // u32 per tile (+1)
    wgb_u64 tile_cursor;                 // u32 per tile
    wgb_u64 bins;                        // u32 entries
    wgb_u64 coverage;                    // optional u32 per pixel (stats)
    wgb_u64 poison;                      // asynchronous submissions: u32 the tile kernels raise / obey (0 = synchronous execution)
    // K0 (wgb_begin_kernel): zeroes `begin_words` words at `counters` (the counters and the per-tile pair counts) and
    // copies the small buffer bindings of a draw that is not waited for to the work set -- snap_src[g][b] (0 = none) to
    // res[g][b].ptr, res[g][b].a bytes -- so that the pass reads them as they were when it was queued, whatever
    // queue.write_buffer puts there while its tile kernel runs
    wgb_u32 begin_words, pad4;
    wgb_u64 snap_src[WGB_MAX_GROUPS][WGB_MAX_BINDINGS];
    wgb_u32 tmap_color, tmap_depth;      // 1: map_color / map_depth describe colour attachment 0 / the depth attachment
    WgbTensorMap map_color, map_depth;
};
