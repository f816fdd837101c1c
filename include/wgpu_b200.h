/*
 * wgpu_b200.h -- C ABI of the B200-native render-pass backend.
 *
 * This is the drop-in boundary for wgpu-cpu's render-pass draw path: every entry point below
 * is what one method of wgpu-cpu's `impl wgpu::custom::*Interface for ...` blocks would bind
 * (the Rust host crate becomes a thin caller), and each cites the
 * reference interface it replaces (paths relative to the reference checkout).
 *
 * Conventions
 *   - objects are opaque, reference counted handles (mirrors the reference's `Clone` over `Arc`,
 *     buffer.rs:22-29, pipeline.rs:51-54): wgb_retain / wgb_release; command buffers keep the
 *     resources they use alive until they have executed;
 *   - every call returns a wgb_status (0 = ok) and never unwinds or aborts across the boundary;
 *     wgb_last_error() returns a thread-local description of the last failure.  Where the
 *     reference `todo!()`s or panics the call returns WGB_ERROR_UNSUPPORTED / WGB_ERROR_VALIDATION;
 *   - all calls are thread safe; submissions execute in submission order on one CUDA stream per
 *     device (engine.rs:26-36 executes them in order on one thread);
 *   - plain pointers and sizes only.  There is no CPU fallback: without a CUDA device
 *     wgb_adapter_request_device fails.
 */
#ifndef WGPU_B200_H
#define WGPU_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WGB_API __attribute__((visibility("default")))

typedef int32_t wgb_status;
enum {
    WGB_OK = 0,
    WGB_ERROR_VALIDATION = 1,    /* the reference would panic / assert */
    WGB_ERROR_UNSUPPORTED = 2,   /* the reference `todo!()`s, or a state combination this backend does not run */
    WGB_ERROR_OUT_OF_MEMORY = 3,
    WGB_ERROR_DEVICE = 4,        /* CUDA / NVRTC failure (no device, compile error, launch error) */
    WGB_ERROR_SHADER = 5,        /* WGSL translation or CUDA compilation of a shader failed */
    WGB_ERROR_OUT_OF_BOUNDS = 6, /* an index / vertex fetch left its buffer (the reference panics on the slice) */
    WGB_ERROR_TIMEOUT = 7
};

typedef struct wgb_object_t* wgb_object; /* any handle */
typedef struct wgb_instance_t* wgb_instance;
typedef struct wgb_adapter_t* wgb_adapter;
typedef struct wgb_device_t* wgb_device;
typedef struct wgb_queue_t* wgb_queue;
typedef struct wgb_buffer_t* wgb_buffer;
typedef struct wgb_texture_t* wgb_texture;
typedef struct wgb_texture_view_t* wgb_texture_view;
typedef struct wgb_sampler_t* wgb_sampler;
typedef struct wgb_shader_module_t* wgb_shader_module;
typedef struct wgb_bind_group_layout_t* wgb_bind_group_layout;
typedef struct wgb_pipeline_layout_t* wgb_pipeline_layout;
typedef struct wgb_bind_group_t* wgb_bind_group;
typedef struct wgb_render_pipeline_t* wgb_render_pipeline;
typedef struct wgb_command_encoder_t* wgb_command_encoder;
typedef struct wgb_render_pass_t* wgb_render_pass;
typedef struct wgb_command_buffer_t* wgb_command_buffer;
typedef struct wgb_surface_t* wgb_surface;

/* ---- enumerations (names follow wgpu-types) ---- */
enum { WGB_TOPOLOGY_POINT_LIST = 0, WGB_TOPOLOGY_LINE_LIST = 1, WGB_TOPOLOGY_LINE_STRIP = 2,
       WGB_TOPOLOGY_TRIANGLE_LIST = 3, WGB_TOPOLOGY_TRIANGLE_STRIP = 4 };
enum { WGB_INDEX_FORMAT_NONE = 0, WGB_INDEX_FORMAT_UINT16 = 1, WGB_INDEX_FORMAT_UINT32 = 2 };
enum { WGB_FRONT_FACE_CCW = 0, WGB_FRONT_FACE_CW = 1 };
enum { WGB_CULL_MODE_NONE = 0, WGB_CULL_MODE_FRONT = 1, WGB_CULL_MODE_BACK = 2 };
enum { WGB_POLYGON_MODE_FILL = 0, WGB_POLYGON_MODE_LINE = 1, WGB_POLYGON_MODE_POINT = 2 };
enum { WGB_COMPARE_NEVER = 1, WGB_COMPARE_LESS = 2, WGB_COMPARE_EQUAL = 3, WGB_COMPARE_LESS_EQUAL = 4,
       WGB_COMPARE_GREATER = 5, WGB_COMPARE_NOT_EQUAL = 6, WGB_COMPARE_GREATER_EQUAL = 7, WGB_COMPARE_ALWAYS = 8 };
enum { WGB_TEXTURE_FORMAT_RGBA8_UNORM = 0, WGB_TEXTURE_FORMAT_RGBA8_UNORM_SRGB = 1,
       WGB_TEXTURE_FORMAT_BGRA8_UNORM = 2, WGB_TEXTURE_FORMAT_BGRA8_UNORM_SRGB = 3,
       WGB_TEXTURE_FORMAT_R8_UNORM = 4, WGB_TEXTURE_FORMAT_RG8_UNORM = 5,
       WGB_TEXTURE_FORMAT_RGBA8_SNORM = 6, WGB_TEXTURE_FORMAT_DEPTH32_FLOAT = 7 };
enum { WGB_ADDRESS_MODE_CLAMP_TO_EDGE = 0, WGB_ADDRESS_MODE_REPEAT = 1, WGB_ADDRESS_MODE_MIRROR_REPEAT = 2,
       WGB_ADDRESS_MODE_CLAMP_TO_BORDER = 3 };
enum { WGB_FILTER_MODE_NEAREST = 0, WGB_FILTER_MODE_LINEAR = 1 };
enum { WGB_VERTEX_STEP_MODE_VERTEX = 0, WGB_VERTEX_STEP_MODE_INSTANCE = 1 };
enum { WGB_VERTEX_FORMAT_FLOAT32 = 0, WGB_VERTEX_FORMAT_FLOAT32X2 = 1, WGB_VERTEX_FORMAT_FLOAT32X3 = 2,
       WGB_VERTEX_FORMAT_FLOAT32X4 = 3, WGB_VERTEX_FORMAT_UINT32 = 4, WGB_VERTEX_FORMAT_SINT32 = 5,
       WGB_VERTEX_FORMAT_UINT32X2 = 6, WGB_VERTEX_FORMAT_UINT32X3 = 7, WGB_VERTEX_FORMAT_UINT32X4 = 8,
       WGB_VERTEX_FORMAT_SINT32X2 = 9, WGB_VERTEX_FORMAT_SINT32X3 = 10, WGB_VERTEX_FORMAT_SINT32X4 = 11 };
enum { WGB_LOAD_OP_CLEAR = 0, WGB_LOAD_OP_LOAD = 1 };
enum { WGB_STORE_OP_STORE = 0, WGB_STORE_OP_DISCARD = 1 };
enum { WGB_SHADER_STAGE_VERTEX = 1, WGB_SHADER_STAGE_FRAGMENT = 2 };
enum { WGB_BINDING_BUFFER = 1, WGB_BINDING_TEXTURE_VIEW = 2, WGB_BINDING_SAMPLER = 3 };
enum { WGB_MAP_MODE_READ = 1, WGB_MAP_MODE_WRITE = 2 };
/* wgpu::BufferUsages bits */
enum { WGB_BUFFER_USAGE_MAP_READ = 1, WGB_BUFFER_USAGE_MAP_WRITE = 2, WGB_BUFFER_USAGE_COPY_SRC = 4,
       WGB_BUFFER_USAGE_COPY_DST = 8, WGB_BUFFER_USAGE_INDEX = 16, WGB_BUFFER_USAGE_VERTEX = 32,
       WGB_BUFFER_USAGE_UNIFORM = 64, WGB_BUFFER_USAGE_STORAGE = 128 };
enum { WGB_POLL_OK = 0, WGB_POLL_QUEUE_EMPTY = 1, WGB_POLL_TIMEOUT = 2 };
/* wgpu::TextureUsages bits */
enum { WGB_TEXTURE_USAGE_COPY_SRC = 1, WGB_TEXTURE_USAGE_COPY_DST = 2, WGB_TEXTURE_USAGE_TEXTURE_BINDING = 4,
       WGB_TEXTURE_USAGE_STORAGE_BINDING = 8, WGB_TEXTURE_USAGE_RENDER_ATTACHMENT = 16 };
enum { WGB_PRESENT_MODE_AUTO_VSYNC = 0, WGB_PRESENT_MODE_AUTO_NO_VSYNC = 1, WGB_PRESENT_MODE_FIFO = 2,
       WGB_PRESENT_MODE_FIFO_RELAXED = 3, WGB_PRESENT_MODE_IMMEDIATE = 4, WGB_PRESENT_MODE_MAILBOX = 5 };
enum { WGB_COMPOSITE_ALPHA_MODE_AUTO = 0, WGB_COMPOSITE_ALPHA_MODE_OPAQUE = 1 };
enum { WGB_SURFACE_STATUS_GOOD = 0 };
#define WGB_WHOLE_SIZE UINT64_MAX
#define WGB_SUBMISSION_ANY UINT64_MAX

/* ---- error reporting ---- */
WGB_API const char* wgb_last_error(void);
WGB_API void wgb_retain(wgb_object obj);
WGB_API void wgb_release(wgb_object obj);
/* version / build information: "wgpu-b200 <n> sm_100a" */
WGB_API const char* wgb_version(void);

/* ---- instance / adapter / device ---- */
/* replaces wgpu_cpu::instance(Config) (wgpu-cpu/src/lib.rs:22-27), InstanceConfig (instance.rs:23-26) */
typedef struct { uint32_t reserved; } wgb_instance_config;
WGB_API wgb_status wgb_create_instance(const wgb_instance_config* config, wgb_instance* out);
/* InstanceInterface::request_adapter / enumerate_adapters (instance.rs:72-110) */
WGB_API wgb_status wgb_instance_request_adapter(wgb_instance instance, wgb_adapter* out);
/* AdapterInterface::get_info (adapter.rs:60-75): name, device type, backend */
typedef struct {
    char name[128];          /* "wgpu-b200 (<CUDA device name>)" */
    uint32_t device_type;    /* 1 = discrete GPU (wgpu::DeviceType::DiscreteGpu) */
    uint32_t cuda_device_count;
} wgb_adapter_info;
WGB_API wgb_status wgb_adapter_get_info(wgb_adapter adapter, wgb_adapter_info* out);
/* AdapterInterface::request_device -> create_device_and_queue (adapter.rs:24-42, device.rs:40-74).
 * The descriptor adds what the reference's empty Config gains on a multi-GPU box:
 * the CUDA ordinal and this device's share of a sort-first screen partition. */
typedef struct {
    int32_t cuda_device;     /* CUDA ordinal; -1 = the calling thread's current device */
    uint32_t band_rank;      /* this device renders tile-row band `band_rank` ... */
    uint32_t band_count;     /* ... of `band_count` contiguous bands (0 or 1 = whole framebuffer) */
    uint32_t features;       /* WGB_FEATURE_* bits: behaviour beyond the reference, off by default (wgpu's required_features) */
} wgb_device_descriptor;
/* Pipeline features that the reference accepts and ignores.  With none requested the backend is
 * in parity mode: results are the reference's, bit for bit.  Each bit switches one WebGPU behaviour on. */
#define WGB_FEATURE_VIEWPORT_DEPTH_RANGE 1u  /* fragment depth = min_depth + ndc.z * (max_depth - min_depth); the reference keeps ndc.z (raster.rs:141-158) */
#define WGB_FEATURE_COLOR_WRITE_MASK 2u      /* wgb_color_target_state.write_mask is applied (ignored by the reference, fragment.rs:480-485) */
#define WGB_FEATURE_SRGB_ENCODE 4u           /* *Srgb colour targets store sRGB-encoded values (the reference stores them linearly, texture.rs:393-400) */
#define WGB_FEATURE_DYNAMIC_OFFSETS 8u       /* set_bind_group dynamic offsets move buffer bindings (stored and ignored by the reference, state.rs:194-205) */
#define WGB_FEATURE_BLEND 16u                /* wgb_color_target_state blend states and set_blend_constant are applied (ignored by the reference,
                                              * fragment.rs:480-485); triangle topologies only; such pipelines take the ordered tile kernel */
#define WGB_COLOR_WRITE_RED 1u
#define WGB_COLOR_WRITE_GREEN 2u
#define WGB_COLOR_WRITE_BLUE 4u
#define WGB_COLOR_WRITE_ALPHA 8u
#define WGB_COLOR_WRITE_ALL 15u
WGB_API wgb_status wgb_adapter_request_device(wgb_adapter adapter, const wgb_device_descriptor* desc,
                                              wgb_device* out_device, wgb_queue* out_queue);
/* DeviceInterface::poll (device.rs:237-295): wait != 0 blocks until `submission_index` (or, for
 * WGB_SUBMISSION_ANY, the next completion) has left the in-flight set; *out_poll is WGB_POLL_*.
 * An error raised while a submission executed (the reference would have panicked on its engine
 * thread) is returned here. */
WGB_API wgb_status wgb_device_poll(wgb_device device, int32_t wait, uint64_t submission_index, uint64_t timeout_ns,
                                   int32_t* out_poll);

/* ---- buffers ---- */
/* DeviceInterface::create_buffer (device.rs:150-160) + Buffer (buffer.rs:23-109).  The reference always
 * creates buffers mapped for writing regardless of mapped_at_creation (buffer.rs:36-38); here the flag is honoured. */
typedef struct { uint64_t size; uint32_t usage; uint32_t mapped_at_creation; } wgb_buffer_descriptor;
WGB_API wgb_status wgb_device_create_buffer(wgb_device device, const wgb_buffer_descriptor* desc, wgb_buffer* out);
/* BufferInterface::map_async (buffer.rs:112-135): completes before returning; callback may be NULL */
typedef void (*wgb_map_callback)(wgb_status status, void* userdata);
WGB_API wgb_status wgb_buffer_map_async(wgb_buffer buffer, uint32_t mode, uint64_t offset, uint64_t size,
                                        wgb_map_callback callback, void* userdata);
/* BufferInterface::get_mapped_range (buffer.rs:137-160): host pointer valid until unmap */
WGB_API wgb_status wgb_buffer_get_mapped_range(wgb_buffer buffer, uint64_t offset, uint64_t size, void** out_ptr);
/* BufferInterface::unmap (buffer.rs:162-172): uploads a write mapping to the device */
WGB_API wgb_status wgb_buffer_unmap(wgb_buffer buffer);
/* QueueInterface::write_buffer (device.rs:332-344): immediate, ordered after earlier submissions; `data` may be reused or
 * freed as soon as the call returns, whatever kind of host memory it is */
WGB_API wgb_status wgb_queue_write_buffer(wgb_queue queue, wgb_buffer buffer, uint64_t offset, const void* data, uint64_t size);
/* The zero-copy variant (no counterpart in the reference, whose buffers are host memory): `data` must be page-locked
 * (cudaHostAlloc / cudaHostRegister) and must stay valid and unchanged until wgb_queue_wait_uploads returns, or until
 * wgb_device_poll has waited for a later submission that uses the buffer.  The copy runs on the device's copy stream,
 * ordered after the last use of this buffer only, so it overlaps rendering that reads other buffers. */
WGB_API wgb_status wgb_queue_write_buffer_pinned_async(wgb_queue queue, wgb_buffer buffer, uint64_t offset, const void* data, uint64_t size);
/* waits for every upload started with wgb_queue_write_buffer_pinned_async */
WGB_API wgb_status wgb_queue_wait_uploads(wgb_queue queue);
/* device address and size of the buffer's storage (zero-copy interop with CUDA libraries, e.g. an NCCL all-gather of a
 * scene that the ranks of a multi-GPU run upload in slices); work on it must be ordered by the caller */
WGB_API wgb_status wgb_buffer_device_pointer(wgb_buffer buffer, uint64_t* out_ptr, uint64_t* out_size);

/* ---- textures / samplers ---- */
/* DeviceInterface::create_texture (device.rs:162-175), Texture::new (texture.rs:28-50) */
typedef struct {
    uint32_t width, height, depth_or_array_layers;
    uint32_t mip_level_count, sample_count;
    uint32_t format;
    uint32_t usage;
} wgb_texture_descriptor;
WGB_API wgb_status wgb_device_create_texture(wgb_device device, const wgb_texture_descriptor* desc, wgb_texture* out);
/* TextureInterface::create_view (texture.rs:52-75); desc may be NULL (default view) */
typedef struct { uint32_t base_array_layer; uint32_t reserved; } wgb_texture_view_descriptor;
WGB_API wgb_status wgb_texture_create_view(wgb_texture texture, const wgb_texture_view_descriptor* desc, wgb_texture_view* out);
/* QueueInterface::write_texture (device.rs:371-434): mip 0, all aspects.  Copies `height` rows of
 * width*bytes_per_texel bytes from rows `bytes_per_row` apart (0 = tightly packed) to origin (x, y). */
WGB_API wgb_status wgb_queue_write_texture(wgb_queue queue, wgb_texture texture, uint32_t x, uint32_t y,
                                           const void* data, uint64_t data_size, uint32_t bytes_per_row,
                                           uint32_t width, uint32_t height);
/* Read-back of a texture's texels, row-major and tightly packed: what wgpu_cpu::dump_texture /
 * image::rgba_texture_image observe (lib.rs:111-173).  Waits for earlier submissions. */
WGB_API wgb_status wgb_texture_read(wgb_texture texture, void* dst, uint64_t dst_size);
/* The same read-back, not waited for (no counterpart in the reference, whose textures are host memory): the texels as
 * the submissions made so far leave them, copied to page-locked `dst` on the device's read-back stream.  Later
 * submissions run while the copy does, and wait for it only where they write this texture; `dst` holds the texels once
 * wgb_device_wait_readbacks has returned. */
WGB_API wgb_status wgb_texture_read_pinned_async(wgb_texture texture, void* dst, uint64_t dst_size);
/* waits for every read-back started with wgb_texture_read_pinned_async */
WGB_API wgb_status wgb_device_wait_readbacks(wgb_device device);
/* wgpu_cpu::dump_texture (lib.rs:111-158): layer 0 of the texture as a PNG file.  Rgba8Unorm[Srgb] bytes as they are,
 * Bgra8Unorm[Srgb] swizzled to RGBA, Depth32Float as 8-bit grey `(depth * 255.0) as u8`; other formats are
 * WGB_ERROR_UNSUPPORTED (`todo!()` in the reference).  The file is a plain (stored, uncompressed) PNG. */
WGB_API wgb_status wgb_texture_dump_png(wgb_texture texture, const char* path);
/* the PNG writer behind it: `channels` = 1 (grey), 3 (RGB) or 4 (RGBA), 8 bits each, rows top to bottom */
WGB_API wgb_status wgb_write_png(const char* path, const void* pixels, uint32_t width, uint32_t height, uint32_t channels);
/* Peer-memory presenter (one process per GPU on an NVLink box): the presenting rank exports its colour target,
 * every other rank imports it and uses the imported texture as ITS colour attachment, so that its tile kernel
 * stores the finished tiles of its band straight into the presenter's memory over NVLink -- compute and the
 * band exchange are one kernel, there is no gather step.  `handle` is a cudaIpcMemHandle_t (64 bytes). */
#define WGB_IPC_HANDLE_SIZE 64
WGB_API wgb_status wgb_texture_export_ipc(wgb_texture texture, uint8_t handle[WGB_IPC_HANDLE_SIZE]);
WGB_API wgb_status wgb_device_import_texture_ipc(wgb_device device, const uint8_t handle[WGB_IPC_HANDLE_SIZE],
                                                 const wgb_texture_descriptor* desc, wgb_texture* out);
/* device address of the texel storage (linear, row-major), for collectives over NVLink */
WGB_API wgb_status wgb_texture_device_pointer(wgb_texture texture, uint64_t* out_ptr, uint64_t* out_size);
/* ---- surface / present (surface.rs) ----
 * The reference presents through softbuffer: a window's CPU pixel buffer, one 32-bit word per pixel, that `present` fills
 * with the surface texture's bytes as they are and hands to the window system (surface.rs:168-193).  On a headless GPU box
 * the window is a *host pixel sink*: page-locked host memory of width*height*4 bytes that `present` fills with one
 * device-to-host copy of the surface texture, and a callback that stands where softbuffer's `Buffer::present` does (a
 * caller forwards the pixels to whatever displays them -- an encoder, a socket, softbuffer itself in the Rust host).
 * Same contract as the reference otherwise: one format (Bgra8Unorm, as softbuffer's 0RGB words), present mode Immediate,
 * alpha Opaque, usage RENDER_ATTACHMENT; one texture per configuration, handed out by every get_current_texture. */
typedef void (*wgb_present_callback)(void* user_data, const void* pixels, uint32_t width, uint32_t height, uint32_t bytes_per_row);
typedef struct { wgb_present_callback on_present; void* user_data; } wgb_surface_target;
/* InstanceInterface::create_surface (instance.rs:51-70), Surface::new (surface.rs:29-49).  `target` may be null or hold a
 * null callback: the presented pixels are then only kept in the window buffer (wgb_surface_get_window_buffer). */
WGB_API wgb_status wgb_instance_create_surface(wgb_instance instance, const wgb_surface_target* target, wgb_surface* out);
/* AdapterInterface::is_surface_supported (adapter.rs:44-54): true for every surface of this library */
WGB_API wgb_status wgb_adapter_is_surface_supported(wgb_adapter adapter, wgb_surface surface, int32_t* out_supported);
typedef struct {
    uint32_t format_count, formats[4];                 /* Bgra8Unorm */
    uint32_t present_mode_count, present_modes[4];     /* Immediate */
    uint32_t alpha_mode_count, alpha_modes[4];         /* Opaque */
    uint32_t usages;                                   /* RENDER_ATTACHMENT */
} wgb_surface_capabilities;
/* SurfaceInterface::get_capabilities (surface.rs:52-72) */
WGB_API wgb_status wgb_surface_get_capabilities(wgb_surface surface, wgb_adapter adapter, wgb_surface_capabilities* out);
typedef struct {
    uint32_t usage, format, width, height, present_mode, alpha_mode;
    uint32_t view_format_count; const uint32_t* view_formats;
} wgb_surface_configuration;
/* SurfaceInterface::configure (surface.rs:74-114): the format and every view format must be Bgra8Unorm
 * (check_surface_config, surface.rs:234-257: the reference unwraps the error) and the extent non-zero (`NonZero::new(..)
 * .expect`), else WGB_ERROR_VALIDATION; allocates the surface texture (zeroed) and sizes the window buffer. */
WGB_API wgb_status wgb_surface_configure(wgb_surface surface, wgb_device device, const wgb_surface_configuration* config);
/* SurfaceInterface::get_current_texture (surface.rs:116-146): a new handle (release it) to the configuration's one
 * texture, status Good; WGB_ERROR_VALIDATION before the first configure ("Surface not configured yet"). */
WGB_API wgb_status wgb_surface_get_current_texture(wgb_surface surface, wgb_texture* out, uint32_t* out_status);
/* SurfaceOutputDetailInterface::present (surface.rs:168-193): waits for the submissions that write the texture (the
 * reference waits for the texture's write guard), copies its bytes into the window buffer and calls `on_present`. */
WGB_API wgb_status wgb_surface_present(wgb_surface surface);
/* SurfaceOutputDetailInterface::texture_discard (surface.rs:195-197): nothing, as in the reference */
WGB_API wgb_status wgb_surface_texture_discard(wgb_surface surface);
/* the window's pixels as the last present left them (valid until the next configure / release), and how many presents
 * this surface has seen */
WGB_API wgb_status wgb_surface_get_window_buffer(wgb_surface surface, const void** out_pixels, uint64_t* out_size, uint64_t* out_presents);
/* DeviceInterface::create_sampler (device.rs:177-180), Sampler (sampler.rs:4-27) */
typedef struct {
    uint32_t address_mode_u, address_mode_v, address_mode_w;
    uint32_t mag_filter, min_filter, mipmap_filter;
} wgb_sampler_descriptor;
WGB_API wgb_status wgb_device_create_sampler(wgb_device device, const wgb_sampler_descriptor* desc, wgb_sampler* out);

/* ---- shaders ---- */
/* DeviceInterface::create_shader_module (device.rs:88-100): WGSL source.  Entry points are translated to
 * CUDA C++ at pipeline creation (the replacement of naga-cranelift's compile_jit, lib.rs:83-113).  A caller
 * that already holds emitter output for an entry point (the Rust host's naga-IR emitter) passes it in
 * `emitted` and the WGSL front end is skipped for that entry point. */
typedef struct { uint32_t stage; const char* entry_point; const char* cuda_source; } wgb_emitted_entry_point;
typedef struct {
    const char* wgsl;                          /* may be NULL if every used entry point is in `emitted` */
    uint32_t emitted_count;
    const wgb_emitted_entry_point* emitted;
} wgb_shader_module_descriptor;
WGB_API wgb_status wgb_device_create_shader_module(wgb_device device, const wgb_shader_module_descriptor* desc,
                                                   wgb_shader_module* out);
/* the WGSL -> CUDA C++ emitter on its own (no device needed); *out_cuda is malloc'ed, free with wgb_free */
WGB_API wgb_status wgb_translate_wgsl(const char* wgsl, uint32_t stage, const char* entry_point, char** out_cuda);
WGB_API void wgb_free(void* p);

/* ---- binding model ---- */
/* DeviceInterface::create_bind_group_layout / create_pipeline_layout (device.rs:102-127): descriptors are kept, not interpreted */
typedef struct { uint32_t binding; uint32_t visibility; uint32_t kind; uint32_t has_dynamic_offset; } wgb_bind_group_layout_entry;
WGB_API wgb_status wgb_device_create_bind_group_layout(wgb_device device, const wgb_bind_group_layout_entry* entries,
                                                       uint32_t count, wgb_bind_group_layout* out);
WGB_API wgb_status wgb_device_create_pipeline_layout(wgb_device device, const wgb_bind_group_layout* layouts,
                                                     uint32_t count, wgb_pipeline_layout* out);
/* DeviceInterface::create_bind_group (device.rs:109-116), BindGroup (bind_group.rs:52-117) */
typedef struct {
    uint32_t binding;
    uint32_t kind;                 /* WGB_BINDING_* */
    wgb_buffer buffer; uint64_t offset; uint64_t size;   /* size WGB_WHOLE_SIZE = to the end */
    wgb_texture_view texture_view;
    wgb_sampler sampler;
} wgb_bind_group_entry;
WGB_API wgb_status wgb_device_create_bind_group(wgb_device device, wgb_bind_group_layout layout,
                                                const wgb_bind_group_entry* entries, uint32_t count, wgb_bind_group* out);

/* ---- render pipeline ---- */
typedef struct { uint32_t format; uint64_t offset; uint32_t shader_location; } wgb_vertex_attribute;
typedef struct { uint64_t array_stride; uint32_t step_mode; uint32_t attribute_count; const wgb_vertex_attribute* attributes; } wgb_vertex_buffer_layout;
/* wgpu::BlendFactor / BlendOperation / BlendComponent */
enum { WGB_BLEND_FACTOR_ZERO = 0, WGB_BLEND_FACTOR_ONE = 1, WGB_BLEND_FACTOR_SRC = 2, WGB_BLEND_FACTOR_ONE_MINUS_SRC = 3,
       WGB_BLEND_FACTOR_SRC_ALPHA = 4, WGB_BLEND_FACTOR_ONE_MINUS_SRC_ALPHA = 5, WGB_BLEND_FACTOR_DST = 6,
       WGB_BLEND_FACTOR_ONE_MINUS_DST = 7, WGB_BLEND_FACTOR_DST_ALPHA = 8, WGB_BLEND_FACTOR_ONE_MINUS_DST_ALPHA = 9,
       WGB_BLEND_FACTOR_SRC_ALPHA_SATURATED = 10, WGB_BLEND_FACTOR_CONSTANT = 11, WGB_BLEND_FACTOR_ONE_MINUS_CONSTANT = 12 };
enum { WGB_BLEND_OPERATION_ADD = 0, WGB_BLEND_OPERATION_SUBTRACT = 1, WGB_BLEND_OPERATION_REVERSE_SUBTRACT = 2,
       WGB_BLEND_OPERATION_MIN = 3, WGB_BLEND_OPERATION_MAX = 4 };
typedef struct { uint32_t src_factor, dst_factor, operation; } wgb_blend_component;
typedef struct { uint32_t format; uint32_t has_blend; uint32_t write_mask; wgb_blend_component blend_color, blend_alpha; } wgb_color_target_state;
typedef struct {
    wgb_pipeline_layout layout;                /* may be NULL */
    /* VertexState (render_pass/vertex.rs:43-93) */
    wgb_shader_module vertex_module; const char* vertex_entry_point;
    uint32_t vertex_buffer_count; const wgb_vertex_buffer_layout* vertex_buffers;
    /* PrimitiveState (state.rs:433-478) */
    uint32_t topology, strip_index_format, front_face, cull_mode, polygon_mode;
    uint32_t unclipped_depth, conservative;
    /* DepthStencilState (fragment.rs:427-453); has_depth_stencil = 0 -> no depth test */
    uint32_t has_depth_stencil, depth_format, depth_write_enabled, depth_compare;
    uint32_t multisample_count;
    /* FragmentState (render_pass/fragment.rs:61-89); fragment_module NULL -> vertex stage only (state.rs:583-588) */
    wgb_shader_module fragment_module; const char* fragment_entry_point;
    uint32_t target_count; const wgb_color_target_state* targets;
} wgb_render_pipeline_descriptor;
/* DeviceInterface::create_render_pipeline (device.rs:129-134) -> RenderPipeline::new (pipeline.rs:57-76):
 * translates both entry points and compiles the pipeline's kernels for sm_100a with NVRTC. */
WGB_API wgb_status wgb_device_create_render_pipeline(wgb_device device, const wgb_render_pipeline_descriptor* desc,
                                                     wgb_render_pipeline* out);
/* the generated CUDA translation unit of the pipeline (diagnostics; *out is malloc'ed) */
WGB_API wgb_status wgb_render_pipeline_get_source(wgb_render_pipeline pipeline, char** out);

/* ---- command encoding ---- */
/* DeviceInterface::create_command_encoder (device.rs:190-196), CommandEncoder (command.rs:11-171) */
WGB_API wgb_status wgb_device_create_command_encoder(wgb_device device, wgb_command_encoder* out);
typedef struct { wgb_texture_view view; uint32_t load_op, store_op; double clear_value[4]; } wgb_color_attachment;
typedef struct { wgb_texture_view view; uint32_t has_depth_ops, depth_load_op, depth_store_op; float depth_clear_value;
                 uint32_t has_stencil_ops; } wgb_depth_stencil_attachment;
typedef struct {
    uint32_t color_attachment_count; const wgb_color_attachment* color_attachments;   /* view NULL = empty slot */
    const wgb_depth_stencil_attachment* depth_stencil_attachment;                       /* may be NULL */
} wgb_render_pass_descriptor;
/* CommandEncoderInterface::begin_render_pass (command.rs:82-87) -> RenderPassEncoder::new (render_pass/mod.rs:53-70) */
WGB_API wgb_status wgb_command_encoder_begin_render_pass(wgb_command_encoder encoder, const wgb_render_pass_descriptor* desc,
                                                         wgb_render_pass* out);
/* RenderPassInterface (render_pass/mod.rs:73-176): every call records one RenderPassSubCommand */
WGB_API wgb_status wgb_render_pass_set_pipeline(wgb_render_pass pass, wgb_render_pipeline pipeline);
WGB_API wgb_status wgb_render_pass_set_bind_group(wgb_render_pass pass, uint32_t index, wgb_bind_group group,
                                                  const uint32_t* dynamic_offsets, uint32_t dynamic_offset_count);
WGB_API wgb_status wgb_render_pass_set_index_buffer(wgb_render_pass pass, wgb_buffer buffer, uint32_t index_format,
                                                    uint64_t offset, uint64_t size);
WGB_API wgb_status wgb_render_pass_set_vertex_buffer(wgb_render_pass pass, uint32_t slot, wgb_buffer buffer,
                                                     uint64_t offset, uint64_t size);
WGB_API wgb_status wgb_render_pass_set_viewport(wgb_render_pass pass, float x, float y, float width, float height,
                                                float min_depth, float max_depth);
WGB_API wgb_status wgb_render_pass_set_scissor_rect(wgb_render_pass pass, uint32_t x, uint32_t y, uint32_t width, uint32_t height);
WGB_API wgb_status wgb_render_pass_set_blend_constant(wgb_render_pass pass, const double color[4]);
WGB_API wgb_status wgb_render_pass_set_stencil_reference(wgb_render_pass pass, uint32_t reference);
WGB_API wgb_status wgb_render_pass_draw(wgb_render_pass pass, uint32_t first_vertex, uint32_t vertex_count,
                                        uint32_t first_instance, uint32_t instance_count);
WGB_API wgb_status wgb_render_pass_draw_indexed(wgb_render_pass pass, uint32_t first_index, uint32_t index_count,
                                                int32_t base_vertex, uint32_t first_instance, uint32_t instance_count);
/* RenderPassInterface::end (render_pass/mod.rs:309-329): pushes Command::RenderPass into the encoder */
WGB_API wgb_status wgb_render_pass_end(wgb_render_pass pass);
/* Encoder copies and clears: CommandEncoderInterface::copy_buffer_to_buffer / copy_buffer_to_texture /
 * copy_texture_to_buffer / copy_texture_to_texture / clear_buffer / clear_texture (command.rs:37-115).  All six are
 * `todo!()` in the reference (its tests read the texture's Vec<u8> directly, lib.rs:111-173); they are implemented
 * here with WebGPU semantics because every standard wgpu app observes results through copy_texture_to_buffer +
 * map_async.  They execute in order with the render passes of the same command buffer.
 * Texel copies address mip 0; `bytes_per_row` is the buffer's row pitch (0 = tightly packed). */
typedef struct { wgb_texture texture; uint32_t x, y, layer; } wgb_texel_copy_texture_info;
typedef struct { wgb_buffer buffer; uint64_t offset; uint32_t bytes_per_row, rows_per_image; } wgb_texel_copy_buffer_info;
WGB_API wgb_status wgb_command_encoder_copy_buffer_to_buffer(wgb_command_encoder encoder, wgb_buffer source, uint64_t source_offset,
                                                             wgb_buffer destination, uint64_t destination_offset, uint64_t size);
WGB_API wgb_status wgb_command_encoder_copy_buffer_to_texture(wgb_command_encoder encoder, const wgb_texel_copy_buffer_info* source,
                                                              const wgb_texel_copy_texture_info* destination, uint32_t width, uint32_t height);
WGB_API wgb_status wgb_command_encoder_copy_texture_to_buffer(wgb_command_encoder encoder, const wgb_texel_copy_texture_info* source,
                                                              const wgb_texel_copy_buffer_info* destination, uint32_t width, uint32_t height);
WGB_API wgb_status wgb_command_encoder_copy_texture_to_texture(wgb_command_encoder encoder, const wgb_texel_copy_texture_info* source,
                                                               const wgb_texel_copy_texture_info* destination, uint32_t width, uint32_t height);
WGB_API wgb_status wgb_command_encoder_clear_buffer(wgb_command_encoder encoder, wgb_buffer buffer, uint64_t offset, uint64_t size);
WGB_API wgb_status wgb_command_encoder_clear_texture(wgb_command_encoder encoder, wgb_texture texture);
/* CommandEncoderInterface::finish (command.rs:89-98) */
WGB_API wgb_status wgb_command_encoder_finish(wgb_command_encoder encoder, wgb_command_buffer* out);
/* QueueInterface::submit (device.rs:436-462): monotonically increasing submission index */
WGB_API wgb_status wgb_queue_submit(wgb_queue queue, const wgb_command_buffer* command_buffers, uint32_t count,
                                    uint64_t* out_submission_index);

/* ---- measurement (mirrors the reference's per-pass Instant timer and per-draw counters,
 *      render_pass/mod.rs:346,392-393 and state.rs:516-517,592) ---- */
typedef struct {
    uint64_t primitives;          /* submitted (assembled) primitives */
    uint64_t fragments;           /* rasterised fragments = the reference's fragment-stage invocations (with coverage capture
                                     on; otherwise fragments of primitives dropped by the hierarchical depth test are not counted) */
    uint64_t shaded;              /* fragment-shader invocations this backend ran for surviving fragments */
    uint64_t bin_pairs;           /* (primitive, tile) pairs in the per-tile bins */
    uint64_t big_primitives;      /* primitives on the all-tiles list */
    uint64_t clipped_primitives;  /* primitives that went through the clipper */
    uint64_t clip_records;
    uint32_t draws, kernel_launches;
    float geometry_ms;            /* geometry + clip + scan + fill kernels (CUDA events) */
    float tile_ms;                /* tile kernels */
    float total_ms;               /* whole pass on the device */
    uint32_t replays;             /* draws replayed after growing a work buffer */
    uint64_t hiz_culled;          /* (primitive, tile) pairs dropped before rasterisation: every fragment provably fails the depth test */
} wgb_pass_stats;
/* statistics of the most recently executed render pass on this device */
WGB_API wgb_status wgb_device_get_last_pass_stats(wgb_device device, wgb_pass_stats* out);
/* parity instrumentation: when enabled, every pass also accumulates a per-pixel count of rasterised
 * fragments (the coverage the reference's rasteriser emits), readable with wgb_device_read_coverage */
WGB_API wgb_status wgb_device_set_coverage_capture(wgb_device device, int32_t enabled);
WGB_API wgb_status wgb_device_read_coverage(wgb_device device, uint32_t* dst, uint64_t pixel_count);
/* change this device's share of the sort-first partition (see wgb_device_descriptor) */
WGB_API wgb_status wgb_device_set_band(wgb_device device, uint32_t band_rank, uint32_t band_count);
/* first and one-past-last pixel row of the band for a framebuffer of `height` rows */
WGB_API wgb_status wgb_device_get_band_rows(wgb_device device, uint32_t height, uint32_t* out_row0, uint32_t* out_row1);
/* CUDA-event timer on the device's stream: begin records an event, end records a second one,
 * waits for it and returns the device time between the two in milliseconds */
WGB_API wgb_status wgb_device_timer_begin(wgb_device device);
WGB_API wgb_status wgb_device_timer_end(wgb_device device, float* out_ms);
/* the CUDA stream submissions run on (cudaStream_t), for interop with collectives */
WGB_API wgb_status wgb_device_get_stream(wgb_device device, void** out_stream);

#ifdef __cplusplus
}
#endif
#endif
