// wgb_api.cpp -- host runtime behind include/wgpu_b200.h.
//
// Mirrors wgpu-cpu's backend objects (instance/adapter/device/queue, buffers, textures, samplers,
// shader modules, bind groups, render pipelines, command encoder / render pass) as thin host-side
// handles whose storage lives in B200 device memory, records render passes exactly like the
// reference's RenderPassEncoder (render_pass/mod.rs:46-322), and executes them as CUDA kernel
// sequences on one stream per device.  All arithmetic of the draw path is in wgb_raster.cuh; the
// kernels are compiled per pipeline with NVRTC for sm_100a.  There is no CPU fallback.
#include "../../include/wgpu_b200.h"
#include "wgb_shared.h"

#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nvrtc.h>
#if defined(__has_include)
#if __has_include(<nvtx3/nvToolsExt.h>)
#include <nvtx3/nvToolsExt.h>
#define WGB_HAVE_NVTX 1
#endif
#endif

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

extern const char* const wgb_embedded_shared_h;
extern const char* const wgb_embedded_prelude_cuh;
extern const char* const wgb_embedded_raster_cuh;
// wgsl_emit.cpp
std::string wgb_emit_wgsl(const std::string& wgsl, uint32_t stage, const std::string& entry_point);

namespace {

// ------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------
thread_local std::string g_last_error;

struct Error : std::runtime_error {
    wgb_status status;
    Error(wgb_status s, const std::string& m) : std::runtime_error(m), status(s) {}
};
[[noreturn]] void fail(wgb_status s, const char* fmt, ...) {
    char buf[4096];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    throw Error(s, buf);
}
#define CUDA_CHECK(expr)                                                                              \
    do {                                                                                              \
        cudaError_t e_ = (expr);                                                                      \
        if (e_ != cudaSuccess) fail(e_ == cudaErrorMemoryAllocation ? WGB_ERROR_OUT_OF_MEMORY : WGB_ERROR_DEVICE, \
                                    "%s failed: %s", #expr, cudaGetErrorString(e_));                  \
    } while (0)

// The emitter reports problems as std::runtime_error("WGSL:<line>: ...").  What WGSL allows and this backend (like the
// reference's JIT for most of it) does not run -- f16, overrides, workgroup variables, texture offsets, ... -- is
// WGB_ERROR_UNSUPPORTED; anything else is a shader error.
std::string emit_wgsl(const std::string& wgsl, uint32_t stage, const std::string& entry_point) {
    try {
        return wgb_emit_wgsl(wgsl, stage, entry_point);
    } catch (const Error&) {
        throw;
    } catch (const std::runtime_error& e) {
        const std::string m = e.what();
        const bool unsupported = m.find("not supported") != std::string::npos || m.find("unsupported") != std::string::npos ||
                                 m.find("only ") != std::string::npos;
        throw Error(unsupported ? WGB_ERROR_UNSUPPORTED : WGB_ERROR_SHADER, m);
    }
}

template <class F>
wgb_status guarded(F f) {
    try {
        f();
        return WGB_OK;
    } catch (const Error& e) {
        g_last_error = e.what();
        return e.status;
    } catch (const std::bad_alloc&) {
        g_last_error = "host allocation failed";
        return WGB_ERROR_OUT_OF_MEMORY;
    } catch (const std::exception& e) {
        g_last_error = e.what();
        return WGB_ERROR_DEVICE;
    } catch (...) {
        g_last_error = "unknown error";
        return WGB_ERROR_DEVICE;
    }
}
#define REQUIRE(cond, ...) do { if (!(cond)) fail(WGB_ERROR_VALIDATION, __VA_ARGS__); } while (0)

// ------------------------------------------------------------------------------------------
// reference-counted objects
// ------------------------------------------------------------------------------------------
struct Object {
    std::atomic<int> rc{1};
    virtual ~Object() {}
};
template <class T>
struct Ref {
    T* p = nullptr;
    Ref() {}
    Ref(T* q) : p(q) { if (p) p->rc.fetch_add(1); }
    Ref(const Ref& o) : p(o.p) { if (p) p->rc.fetch_add(1); }
    Ref(Ref&& o) noexcept : p(o.p) { o.p = nullptr; }
    Ref& operator=(Ref o) { std::swap(p, o.p); return *this; }
    ~Ref() { if (p && p->rc.fetch_sub(1) == 1) delete p; }
    T* operator->() const { return p; }
    T* get() const { return p; }
    explicit operator bool() const { return p != nullptr; }
};
template <class T, class H>
T* from_handle(H h, const char* what) {
    if (!h) fail(WGB_ERROR_VALIDATION, "null %s handle", what);
    T* t = dynamic_cast<T*>(reinterpret_cast<Object*>(h));
    if (!t) fail(WGB_ERROR_VALIDATION, "handle is not a %s", what);
    return t;
}
template <class H, class T>
H to_handle(T* t) { return reinterpret_cast<H>(static_cast<Object*>(t)); }

// ------------------------------------------------------------------------------------------
// CUDA driver entry points (through the statically linked runtime, so the library loads on
// machines without libcuda) and NVRTC (dlopen'ed on first use)
// ------------------------------------------------------------------------------------------
struct DriverApi {
    CUresult (*ModuleLoadData)(CUmodule*, const void*) = nullptr;
    CUresult (*ModuleUnload)(CUmodule) = nullptr;
    CUresult (*ModuleGetFunction)(CUfunction*, CUmodule, const char*) = nullptr;
    CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void**, void**) = nullptr;
    CUresult (*GetErrorString)(CUresult, const char**) = nullptr;
    // optional (CUDA 12): tensor maps for the tile kernel's one-copy tile loads / stores
    CUresult (*TensorMapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                     const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill) = nullptr;
    bool loaded = false;
};
DriverApi g_drv;
std::mutex g_global_mu;

void load_driver_api() {
    std::lock_guard<std::mutex> lk(g_global_mu);
    if (g_drv.loaded) return;
    auto get = [](const char* name, void** fn) {
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &q);
        if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !*fn)
            fail(WGB_ERROR_DEVICE, "cudaGetDriverEntryPoint(%s) failed: %s", name, cudaGetErrorString(e));
    };
    get("cuModuleLoadData", (void**)&g_drv.ModuleLoadData);
    get("cuModuleUnload", (void**)&g_drv.ModuleUnload);
    get("cuModuleGetFunction", (void**)&g_drv.ModuleGetFunction);
    get("cuLaunchKernel", (void**)&g_drv.LaunchKernel);
    get("cuGetErrorString", (void**)&g_drv.GetErrorString);
    {
        cudaDriverEntryPointQueryResult q;
        void* fn = nullptr;
        if (!getenv("WGB_NO_TENSOR_MAP") && cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess && fn)
            g_drv.TensorMapEncodeTiled = (decltype(g_drv.TensorMapEncodeTiled))fn;
        cudaGetLastError();
    }
    g_drv.loaded = true;
}
const char* cu_error(CUresult r) {
    const char* s = nullptr;
    if (g_drv.GetErrorString) g_drv.GetErrorString(r, &s);
    return s ? s : "unknown CUDA driver error";
}

struct NvrtcApi {
    void* lib = nullptr;
    nvrtcResult (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
    nvrtcResult (*DestroyProgram)(nvrtcProgram*) = nullptr;
    nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
    nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t*) = nullptr;
    nvrtcResult (*GetCUBIN)(nvrtcProgram, char*) = nullptr;
    nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
    nvrtcResult (*GetProgramLog)(nvrtcProgram, char*) = nullptr;
    const char* (*GetErrorString)(nvrtcResult) = nullptr;
};
NvrtcApi g_nvrtc;

void load_nvrtc() {
    std::lock_guard<std::mutex> lk(g_global_mu);
    if (g_nvrtc.lib) return;
    const char* names[] = {"libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so",
                           "/usr/local/cuda/lib64/libnvrtc.so"};
    void* lib = nullptr;
    for (const char* n : names) { lib = dlopen(n, RTLD_NOW | RTLD_LOCAL); if (lib) break; }
    if (!lib) fail(WGB_ERROR_DEVICE, "cannot load NVRTC (libnvrtc.so.12): %s", dlerror());
    auto sym = [&](const char* n) { void* p = dlsym(lib, n); if (!p) fail(WGB_ERROR_DEVICE, "NVRTC symbol %s missing", n); return p; };
    g_nvrtc.CreateProgram = (decltype(g_nvrtc.CreateProgram))sym("nvrtcCreateProgram");
    g_nvrtc.DestroyProgram = (decltype(g_nvrtc.DestroyProgram))sym("nvrtcDestroyProgram");
    g_nvrtc.CompileProgram = (decltype(g_nvrtc.CompileProgram))sym("nvrtcCompileProgram");
    g_nvrtc.GetCUBINSize = (decltype(g_nvrtc.GetCUBINSize))sym("nvrtcGetCUBINSize");
    g_nvrtc.GetCUBIN = (decltype(g_nvrtc.GetCUBIN))sym("nvrtcGetCUBIN");
    g_nvrtc.GetProgramLogSize = (decltype(g_nvrtc.GetProgramLogSize))sym("nvrtcGetProgramLogSize");
    g_nvrtc.GetProgramLog = (decltype(g_nvrtc.GetProgramLog))sym("nvrtcGetProgramLog");
    g_nvrtc.GetErrorString = (decltype(g_nvrtc.GetErrorString))sym("nvrtcGetErrorString");
    g_nvrtc.lib = lib;
}

// compile one pipeline translation unit to an sm_100a cubin
std::vector<char> nvrtc_compile(const std::string& source) {
    load_nvrtc();
    const char* header_names[] = {"wgb_shared.h", "wgb_prelude.cuh", "wgb_raster.cuh"};
    const char* headers[] = {wgb_embedded_shared_h, wgb_embedded_prelude_cuh, wgb_embedded_raster_cuh};
    nvrtcProgram prog;
    nvrtcResult r = g_nvrtc.CreateProgram(&prog, source.c_str(), "wgb_pipeline.cu", 3, headers, header_names);
    if (r != NVRTC_SUCCESS) fail(WGB_ERROR_DEVICE, "nvrtcCreateProgram: %s", g_nvrtc.GetErrorString(r));
    // --fmad=false: the reference never fuses a multiply with an add; division and
    // square root stay IEEE (--prec-div / --prec-sqrt default true), no flush-to-zero
    const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "--fmad=false", "-lineinfo", "-w"};
    r = g_nvrtc.CompileProgram(prog, 5, opts);
    if (r != NVRTC_SUCCESS) {
        size_t n = 0;
        g_nvrtc.GetProgramLogSize(prog, &n);
        std::string log(n, '\0');
        if (n) g_nvrtc.GetProgramLog(prog, &log[0]);
        g_nvrtc.DestroyProgram(&prog);
        if (log.size() > 3500) log.resize(3500);
        fail(WGB_ERROR_SHADER, "NVRTC compilation failed: %s\n%s", g_nvrtc.GetErrorString(r), log.c_str());
    }
    size_t n = 0;
    g_nvrtc.GetCUBINSize(prog, &n);
    std::vector<char> cubin(n);
    g_nvrtc.GetCUBIN(prog, cubin.data());
    g_nvrtc.DestroyProgram(&prog);
    return cubin;
}

// ------------------------------------------------------------------------------------------
// formats
// ------------------------------------------------------------------------------------------
// overflow-safe "offset + size <= limit"
bool range_ok(uint64_t offset, uint64_t size, uint64_t limit) { return offset <= limit && size <= limit - offset; }

uint32_t bytes_per_texel(uint32_t format) {          // texture.rs:457-509
    switch (format) {
        case WGB_TEXTURE_FORMAT_R8_UNORM: return 1;
        case WGB_TEXTURE_FORMAT_RG8_UNORM: return 2;
        case WGB_TEXTURE_FORMAT_RGBA8_UNORM: case WGB_TEXTURE_FORMAT_RGBA8_UNORM_SRGB:
        case WGB_TEXTURE_FORMAT_BGRA8_UNORM: case WGB_TEXTURE_FORMAT_BGRA8_UNORM_SRGB:
        case WGB_TEXTURE_FORMAT_RGBA8_SNORM: case WGB_TEXTURE_FORMAT_DEPTH32_FLOAT: return 4;
        default: return 0;
    }
}
bool is_color_format(uint32_t f) { return f <= WGB_TEXTURE_FORMAT_RGBA8_SNORM; }
// TexelWriter::from_color (texture.rs:373-411) for the clear colour, on the host
uint32_t f32_to_u8(float v) {
    float s = v * 255.0f;
    if (!(s == s)) return 0;
    if (s < 0.0f) s = 0.0f;
    if (s > 255.0f) s = 255.0f;
    return (uint32_t)s;
}
bool is_srgb_format(uint32_t f) { return f == WGB_TEXTURE_FORMAT_RGBA8_UNORM_SRGB || f == WGB_TEXTURE_FORMAT_BGRA8_UNORM_SRGB; }
float srgb_oetf(float v) {      // same operation sequence as wgb_srgb_oetf on the device
    if (!(v > 0.0031308f)) return v * 12.92f;
    return 1.055f * powf(v, 1.0f / 2.4f) - 0.055f;
}
// colour write mask (WGB_COLOR_WRITE_* bits) -> byte mask over the stored texel
uint32_t texel_write_mask(uint32_t mask, uint32_t format) {
    const bool bgra = format == WGB_TEXTURE_FORMAT_BGRA8_UNORM || format == WGB_TEXTURE_FORMAT_BGRA8_UNORM_SRGB;
    uint32_t m = 0;
    if (mask & WGB_COLOR_WRITE_RED) m |= bgra ? 0x00FF0000u : 0x000000FFu;
    if (mask & WGB_COLOR_WRITE_GREEN) m |= 0x0000FF00u;
    if (mask & WGB_COLOR_WRITE_BLUE) m |= bgra ? 0x000000FFu : 0x00FF0000u;
    if (mask & WGB_COLOR_WRITE_ALPHA) m |= 0xFF000000u;
    return m;
}
uint32_t encode_color(const double c[4], uint32_t format, bool srgb_encode = false) {
    // wgpu_color_to_vec4 casts the f64 colour to f32 first (texture.rs:505-507)
    float v[4] = {(float)c[0], (float)c[1], (float)c[2], (float)c[3]};
    if (srgb_encode) for (int i = 0; i < 3; i++) v[i] = srgb_oetf(v[i]);
    const uint32_t r = f32_to_u8(v[0]), g = f32_to_u8(v[1]), b = f32_to_u8(v[2]), a = f32_to_u8(v[3]);
    if (format == WGB_TEXTURE_FORMAT_BGRA8_UNORM || format == WGB_TEXTURE_FORMAT_BGRA8_UNORM_SRGB)
        return b | (g << 8) | (r << 16) | (a << 24);
    return r | (g << 8) | (b << 16) | (a << 24);
}
uint32_t vertex_format_size(uint32_t f) {
    switch (f) {
        case WGB_VERTEX_FORMAT_FLOAT32: case WGB_VERTEX_FORMAT_UINT32: case WGB_VERTEX_FORMAT_SINT32: return 4;
        case WGB_VERTEX_FORMAT_FLOAT32X2: case WGB_VERTEX_FORMAT_UINT32X2: case WGB_VERTEX_FORMAT_SINT32X2: return 8;
        case WGB_VERTEX_FORMAT_FLOAT32X3: case WGB_VERTEX_FORMAT_UINT32X3: case WGB_VERTEX_FORMAT_SINT32X3: return 12;
        case WGB_VERTEX_FORMAT_FLOAT32X4: case WGB_VERTEX_FORMAT_UINT32X4: case WGB_VERTEX_FORMAT_SINT32X4: return 16;
        default: return 0;
    }
}

// ------------------------------------------------------------------------------------------
// device memory helper
// ------------------------------------------------------------------------------------------
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    void ensure(size_t bytes) {
        if (bytes <= cap) return;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        const size_t want = bytes + bytes / 4 + 256;
        CUDA_CHECK(cudaMalloc(&p, want));
        cap = want;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    uint64_t addr() const { return (uint64_t)(uintptr_t)p; }
};

// kernels of one compiled pipeline variant
struct KernelSet {
    CUmodule module = nullptr;
    CUfunction begin = nullptr, geometry = nullptr, vertex = nullptr, geometry_cached = nullptr, clip = nullptr, scan = nullptr, fill = nullptr, tile = nullptr, clear = nullptr, strip_map = nullptr,
               big_sort = nullptr;     // ordered variants only
    bool ordered = false;
};

struct Instance : Object {};
struct Adapter : Object { Ref<Instance> instance; };

struct Device : Object {
    int ordinal = -1;
    bool compile_only = false;
    cudaStream_t stream = nullptr;
    // The tile kernels of passes that are not waited for run on their own stream, so that the geometry stage of the next
    // pass (render stream, its own set of work buffers) runs beside the tile kernel of this one: the first is bound by
    // latency and memory, the second by instruction issue.  Everything else the render stream does first waits for the
    // tile stream (joined()); what only has to FOLLOW the queued work -- the events that mark a submission done or a
    // buffer used -- is recorded behind both (tail()).  WGB_NO_OVERLAP=1 keeps the tile kernels on the render stream.
    cudaStream_t tile_stream = nullptr;
    cudaEvent_t ev_tiles = nullptr, ev_render_pos = nullptr;
    bool tiles_ahead = false;              // the tile stream holds work the render stream has not waited for
    bool overlap_stages = true;
    // The tile stream pays only where the next geometry stage follows without anything else on the render stream in
    // between (a queue.write_buffer of the next frame's camera makes the render stream wait for the tile kernel, which
    // reads the old one; the two cross-stream waits then cost more than nothing at all).  A draw takes the tile stream
    // when the draw before it was followed by nothing but this one.
    uint32_t ops_since_tile = 1;
    // Which writes have to wait for tile kernels: a draw that is not waited for reads its small buffer bindings from a
    // snapshot in its work set (K0), its vertex / index buffers and large bindings where they are.  The draws are
    // numbered; a buffer remembers the last draw that reads it in place, the device the last draw the render stream has
    // been ordered behind.  A small queue.write_buffer goes behind the tile stream only if its buffer is read in place
    // by a draw after that one.
    uint64_t draw_serial = 0, joined_draw_serial = 0;
    cudaStream_t joined() {
        ops_since_tile++;
        if (tiles_ahead) {
            cudaEventRecord(ev_tiles, tile_stream);
            cudaStreamWaitEvent(stream, ev_tiles, 0);
            tiles_ahead = false;
        }
        joined_draw_serial = draw_serial;
        return stream;
    }
    cudaStream_t tail() {
        if (!tiles_ahead) return stream;
        cudaEventRecord(ev_render_pos, stream);
        cudaStreamWaitEvent(tile_stream, ev_render_pos, 0);
        return tile_stream;
    }
    cudaStream_t copy_stream = nullptr;    // large host->device uploads (wgb_queue_write_buffer from pinned memory)
    cudaStream_t readback_stream = nullptr;    // device->host read-backs that are not waited for (wgb_texture_read_pinned_async)
    cudaEvent_t ev_readback_order = nullptr;
    std::recursive_mutex mu;
    uint32_t band_rank = 0, band_count = 1;
    uint32_t features = 0;                  // WGB_FEATURE_* (opt-in behaviour beyond the reference)
    // submissions (device.rs:436-462, 517-541)
    uint64_t next_submission = 1;
    struct InFlight { uint64_t index; cudaEvent_t done; };
    std::deque<InFlight> inflight;
    wgb_status deferred_status = WGB_OK;   // error raised while executing a submission
    std::string deferred_error;
    // work buffers, grown on demand and reused across passes
    std::map<std::pair<uint32_t, uint32_t>, uint32_t> bin_cap_hint;   // (primitives, band tiles) of a draw -> slots per tile it needed
    bool no_direct_bins = false;
    bool small_work_buffers = false;       // testing knob: start the clip-record and big lists at 2 entries so that both overflow-and-replay paths run
    // the buffers a draw's kernels hand to each other; two sets, used in turn by the draws that are not waited for, so
    // that a geometry stage can fill one while the tile kernel of the draw before still reads the other
    struct WorkSet {
        DevBuf counters, prim_box, setup_cache, vcache_raster, vcache_ndc, vcache_flags, slow_list, clip_records, big_list, tile_offset, tile_cursor, bins, snapshots;
        cudaEvent_t ev_free = nullptr;     // on the tile stream, behind the tile kernel that read the set last
        bool busy = false;
        void release() {
            DevBuf* bufs[] = {&counters, &prim_box, &setup_cache, &vcache_raster, &vcache_ndc, &vcache_flags, &slow_list, &clip_records, &big_list, &tile_offset, &tile_cursor, &bins, &snapshots};
            for (DevBuf* b : bufs) b->release();
            if (ev_free) cudaEventDestroy(ev_free);
            ev_free = nullptr;
        }
    };
    WorkSet work[2];
    uint32_t work_next = 0;
    DevBuf coverage, strip_map, strip_count;
    WgbCounters* host_counters = nullptr;   // pinned
    uint32_t clip_capacity = 0, big_capacity = 0;
    bool coverage_capture = false;
    uint32_t coverage_w = 0, coverage_h = 0;
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t timer_ev[2] = {nullptr, nullptr};
    wgb_pass_stats last_stats{};
    std::map<std::string, std::shared_ptr<KernelSet>> kernel_cache;   // by translation-unit text

    // Asynchronous submissions (device.rs:436-462: submit returns an index at once, poll waits).  A render pass whose
    // attachments are all cleared on load is enqueued without waiting for its counters; what it left behind -- counters
    // of every draw batch in pinned memory, events for the timings -- is looked at when somebody waits (`settle`).  If a
    // batch overflowed its work buffers or raised an error, its tile kernel has set the `poison` word and every later
    // tile / clear kernel has left the attachments alone: settle re-runs the failed pass and everything queued behind
    // it, synchronously.  (Passes that load their attachments, strips with primitive restart and the coverage capture
    // of the parity tests are executed synchronously as before; so is everything with WGB_SYNC_SUBMIT=1.)
    bool async_submit = true;
    struct PendingBatch { WgbCounters* host; cudaEvent_t ev[4]; uint32_t np, band_tiles; };      // events: geometry stage begins / ends, tile kernel ends / begins
    struct Unsettled {
        uint64_t submission = 0;
        std::shared_ptr<struct PassCommand> pass;              // a render pass that was enqueued without waiting, or ...
        struct Buffer* write_buffer = nullptr;                 // ... a small queue.write_buffer issued behind one (kept alive by write_ref)
        std::shared_ptr<void> write_ref;
        uint64_t write_offset = 0;
        std::vector<uint8_t> write_data;
        std::vector<PendingBatch> batches;
        wgb_pass_stats stats{};
    };
    std::deque<Unsettled> unsettled;
    // What the small writes made before the first entry of `unsettled` left in their regions (the latest one per region):
    // a pass that is run again from the head of the log has to see those regions as they were then, although writes
    // queued behind it have gone over them since.  A small write is queued behind passes only if its exact region has an
    // entry here; otherwise the passes are settled first, and the write becomes the entry.  Anything else that changes
    // a buffer (large writes, copies, unmap) drops the buffer's entries.
    struct BaseWrite { struct Buffer* buffer; uint64_t offset; std::vector<uint8_t> data; };
    std::deque<BaseWrite> base_writes;
    static constexpr size_t BASE_WRITES_MAX = 64;
    void forget_base_writes(const struct Buffer* b) {
        for (auto it = base_writes.begin(); it != base_writes.end();) it = (it->buffer == b) ? base_writes.erase(it) : it + 1;
    }
    const BaseWrite* find_base_write(const struct Buffer* b, uint64_t offset, uint64_t size) const {
        for (const auto& w : base_writes) if (w.buffer == b && w.offset == offset && w.data.size() == size) return &w;
        return nullptr;
    }
    void note_base_write(struct Buffer* b, uint64_t offset, const void* data, uint64_t size) {
        for (auto it = base_writes.begin(); it != base_writes.end();)      // the regions this write goes over (its own too)
            it = (it->buffer == b && it->offset < offset + size && offset < it->offset + it->data.size()) ? base_writes.erase(it) : it + 1;
        base_writes.push_back(BaseWrite{b, offset, std::vector<uint8_t>((const uint8_t*)data, (const uint8_t*)data + size)});
        if (base_writes.size() > BASE_WRITES_MAX) base_writes.pop_front();
    }
    Unsettled* recording = nullptr;        // the pass being enqueued asynchronously (execute_draw appends its batches)
    bool settling = false;
    uint64_t current_submission = 0;
    DevBuf poison;
    static constexpr uint32_t COUNTER_RING = 2048;
    WgbCounters* counter_ring = nullptr;   // pinned
    uint32_t counter_ring_used = 0;
    std::vector<cudaEvent_t> event_pool;
    size_t event_pool_used = 0;
    cudaEvent_t pool_event() {
        if (event_pool_used == event_pool.size()) { cudaEvent_t e; CUDA_CHECK(cudaEventCreate(&e)); event_pool.push_back(e); }
        return event_pool[event_pool_used++];
    }

    void make_current() const { if (!compile_only) CUDA_CHECK(cudaSetDevice(ordinal)); }
    ~Device() override {
        if (compile_only) return;
        cudaSetDevice(ordinal);
        if (stream) cudaStreamSynchronize(stream);
        if (tile_stream) { cudaStreamSynchronize(tile_stream); cudaStreamDestroy(tile_stream); }
        if (ev_tiles) cudaEventDestroy(ev_tiles);
        if (ev_render_pos) cudaEventDestroy(ev_render_pos);
        if (copy_stream) { cudaStreamSynchronize(copy_stream); cudaStreamDestroy(copy_stream); }
        if (readback_stream) { cudaStreamSynchronize(readback_stream); cudaStreamDestroy(readback_stream); }
        if (ev_readback_order) cudaEventDestroy(ev_readback_order);
        for (auto& f : inflight) cudaEventDestroy(f.done);
        for (auto& kv : kernel_cache) if (kv.second->module && g_drv.ModuleUnload) g_drv.ModuleUnload(kv.second->module);
        for (auto& w : work) w.release();
        DevBuf* bufs[] = {&coverage, &strip_map, &strip_count};
        for (DevBuf* b : bufs) b->release();
        if (host_counters) cudaFreeHost(host_counters);
        if (counter_ring) cudaFreeHost(counter_ring);
        for (auto& e : event_pool) cudaEventDestroy(e);
        poison.release();
        for (auto& e : ev) if (e) cudaEventDestroy(e);
        for (auto& e : timer_ev) if (e) cudaEventDestroy(e);
        if (stream) cudaStreamDestroy(stream);
    }
};
struct Queue : Object { Ref<Device> device; };

struct Buffer : Object {
    Ref<Device> device;
    uint64_t size = 0;
    uint32_t usage = 0;
    void* dptr = nullptr;
    std::mutex mu;
    std::vector<uint8_t> staging;   // host mapping
    bool mapped = false;
    uint32_t map_mode = 0;
    uint64_t map_offset = 0, map_size = 0;
    // hazard tracking between the device's render stream and its copy stream (large write_buffer uploads from
    // pinned memory run on the copy stream so that they overlap rendering that does not touch this buffer);
    // guarded by the device mutex
    cudaEvent_t ev_write = nullptr, ev_use = nullptr;
    bool write_pending = false, use_pending = false;
    uint64_t last_use_submission = 0;       // the latest submission that reads or writes the buffer (guarded by the device mutex)
    uint64_t last_live_draw = 0;            // Device::draw_serial of the last draw that reads the buffer in place
    bool external_writers = false;          // wgb_buffer_device_pointer has handed the storage out: its contents are the caller's business
    ~Buffer() override {
        {
            std::lock_guard<std::recursive_mutex> dl(device->mu);
            device->forget_base_writes(this);
        }
        if (!dptr) return;
        cudaSetDevice(device->ordinal);
        if (write_pending) cudaEventSynchronize(ev_write);
        if (ev_write) cudaEventDestroy(ev_write);
        if (ev_use) cudaEventDestroy(ev_use);
        cudaFree(dptr);
    }
    // make the render stream wait for an upload still in flight on the copy stream
    void acquire_on(cudaStream_t stream) {
        if (write_pending) { cudaStreamWaitEvent(stream, ev_write, 0); write_pending = false; }
    }
    // note that work using this buffer has been enqueued on `stream`
    void mark_used(cudaStream_t stream) {
        if (!ev_use) cudaEventCreateWithFlags(&ev_use, cudaEventDisableTiming);
        cudaEventRecord(ev_use, stream);
        use_pending = true;
    }
};

struct Texture : Object {
    Ref<Device> device;
    wgb_texture_descriptor desc{};
    uint32_t bpp = 0;
    uint64_t size = 0;
    void* dptr = nullptr;
    bool imported = false;          // dptr maps another process's allocation (cudaIpcOpenMemHandle)
    cudaTextureObject_t texobj = 0;
    std::mutex mu;
    // a read-back in flight on the device's read-back stream (wgb_texture_read_pinned_async): work on the render stream
    // that writes the texture waits for it; guarded by the device mutex
    cudaEvent_t ev_read = nullptr;
    bool read_pending = false;
    void acquire_for_write(cudaStream_t stream) {
        if (read_pending) { cudaStreamWaitEvent(stream, ev_read, 0); read_pending = false; }
    }
    ~Texture() override {
        if (device->compile_only) return;
        cudaSetDevice(device->ordinal);
        if (read_pending) cudaEventSynchronize(ev_read);
        if (ev_read) cudaEventDestroy(ev_read);
        if (texobj) cudaDestroyTextureObject(texobj);
        if (dptr) { if (imported) cudaIpcCloseMemHandle(dptr); else cudaFree(dptr); }
    }
    // bindless texture object over the linear texel storage (element fetch, no filtering)
    cudaTextureObject_t texture_object() {
        std::lock_guard<std::mutex> lk(mu);
        if (texobj) return texobj;
        cudaResourceDesc rd;
        memset(&rd, 0, sizeof(rd));
        rd.resType = cudaResourceTypeLinear;
        rd.res.linear.devPtr = dptr;
        rd.res.linear.desc = cudaCreateChannelDesc<uchar4>();
        rd.res.linear.sizeInBytes = size;
        cudaTextureDesc td;
        memset(&td, 0, sizeof(td));
        td.readMode = cudaReadModeElementType;
        td.filterMode = cudaFilterModePoint;
        td.addressMode[0] = cudaAddressModeClamp;
        td.normalizedCoords = 0;
        CUDA_CHECK(cudaCreateTextureObject(&texobj, &rd, &td, nullptr));
        return texobj;
    }
};
struct TextureView : Object { Ref<Texture> texture; uint32_t base_layer = 0; };
// surface.rs:24-27,148-166: the window (here: a page-locked host pixel buffer + a present callback) and the current
// configuration's one texture
struct Surface : Object {
    std::recursive_mutex mu;         // (recursive: the present callback may ask the surface for its next texture)
    wgb_surface_target target{};
    bool configured = false;
    wgb_surface_configuration config{};
    Ref<Device> device;
    Ref<Texture> texture;
    void* window = nullptr;          // cudaMallocHost: width * height * 4 bytes
    uint64_t window_size = 0;
    uint64_t presents = 0;
    bool window_pinned = false;
    void drop_window() {
        if (window && window_pinned) { cudaSetDevice(device->ordinal); cudaFreeHost(window); }
        else if (window) free(window);
        window = nullptr; window_size = 0; window_pinned = false;
    }
    ~Surface() override { drop_window(); }
};
struct Sampler : Object { wgb_sampler_descriptor desc{}; };

struct ShaderModule : Object {
    std::string wgsl;
    struct Emitted { uint32_t stage; std::string entry, cuda; };
    std::vector<Emitted> emitted;
    std::string cuda_for(uint32_t stage, const std::string& entry) const {
        for (const auto& e : emitted)
            if (e.stage == stage && e.entry == entry) return e.cuda;
        if (wgsl.empty()) fail(WGB_ERROR_SHADER, "shader module has neither WGSL nor emitted CUDA for entry point '%s'", entry.c_str());
        return emit_wgsl(wgsl, stage, entry);
    }
};
struct BindGroupLayout : Object { std::vector<wgb_bind_group_layout_entry> entries; };
struct PipelineLayout : Object { std::vector<Ref<BindGroupLayout>> layouts; };
struct BindGroup : Object {
    Ref<BindGroupLayout> layout;       // names the bindings that take a dynamic offset
    struct Entry {
        uint32_t binding, kind;
        Ref<Buffer> buffer; uint64_t offset = 0, size = 0;
        Ref<TextureView> view;
        Ref<Sampler> sampler;
    };
    std::vector<Entry> entries;
};

struct RenderPipeline : Object {
    Ref<Device> device;
    std::string vs_text, fs_text;
    bool has_fragment = false;
    struct VB { uint64_t stride; uint32_t step_mode; std::vector<wgb_vertex_attribute> attrs; };
    std::vector<VB> vbs;
    uint32_t topology = 0, strip_index_format = 0, front_face = 0, cull_mode = 0;
    bool has_depth_stencil = false;
    uint32_t depth_format = 0, depth_write = 0, depth_compare = 0;
    std::vector<wgb_color_target_state> targets;
    std::mutex mu;
    std::map<uint32_t, std::shared_ptr<KernelSet>> variants;   // key: has_depth_attachment | separated << 1
    std::map<uint32_t, std::string> variant_source;

    bool blends() const {       // WGB_FEATURE_BLEND: a target with a blend state sends the pipeline down the ordered path
        if (!(device->features & WGB_FEATURE_BLEND)) return false;
        for (const auto& t : targets) if (t.has_blend) return true;
        return false;
    }
    int prim_size() const {
        return topology == WGB_TOPOLOGY_POINT_LIST ? 1 : (topology == WGB_TOPOLOGY_LINE_LIST || topology == WGB_TOPOLOGY_LINE_STRIP) ? 2 : 3;
    }
    bool is_strip() const { return topology == WGB_TOPOLOGY_LINE_STRIP || topology == WGB_TOPOLOGY_TRIANGLE_STRIP; }

    // closed form of the serial depth-test fold for this state (see wgb_raster.cuh)
    static int resolve_mode(bool test, uint32_t compare, bool write) {
        if (!test) return 0;
        switch (compare) {
            case WGB_COMPARE_NEVER: return 6;
            case WGB_COMPARE_ALWAYS: return 0;
            case WGB_COMPARE_LESS: return write ? 1 : 5;
            case WGB_COMPARE_LESS_EQUAL: return write ? 2 : 5;
            case WGB_COMPARE_GREATER: return write ? 3 : 5;
            case WGB_COMPARE_GREATER_EQUAL: return write ? 4 : 5;
            case WGB_COMPARE_EQUAL: return 5;
            case WGB_COMPARE_NOT_EQUAL: return write ? 7 : 5;    // with writes every outcome depends on the previous fragment: ordered kernel
            default: fail(WGB_ERROR_VALIDATION, "invalid depth compare function %u", compare);
        }
    }

    // value of a `#define NAME n` the emitter wrote into the fragment stage's text (0 if there is no fragment stage)
    int fs_flag(const char* name) const {
        if (!has_fragment) return 0;
        const std::string key = std::string("#define ") + name + " ";
        const size_t at = fs_text.find(key);
        return at == std::string::npos ? 0 : atoi(fs_text.c_str() + at + key.size());
    }

    std::string build_source(bool has_depth_attachment, bool separated) const {
        const bool test = has_depth_stencil && has_depth_attachment;
        int mode = resolve_mode(test, depth_compare, depth_write != 0);
        if (blends()) mode = 7;
        // an early depth test (fragment.rs:166-194) is the late one in disguise unless the shader can discard or write
        // frag_depth after it, or the late test runs as well (Allow): those take the ordered kernel too
        const int early = fs_flag("WGB_FS_EARLY_DEPTH");
        if (test && early != 0 && (early == 2 || fs_flag("WGB_FS_MAY_DISCARD") || fs_flag("WGB_FS_WRITES_FRAG_DEPTH"))) mode = 7;
        // ... and a stage without colour outputs never reaches the late test (it runs at the first @location output,
        // fragment.rs:457-488): its early test is then the only one that writes depth (a depth-only pass)
        if (test && early != 0 && fs_flag("WGB_FS_COLOR_MASK") == 0) mode = 7;
        std::string s;
        char line[256];
        auto def = [&](const char* name, long long v) { snprintf(line, sizeof(line), "#define %s %lld\n", name, v); s += line; };
        s += "// generated by wgpu-b200: one translation unit per render pipeline variant\n";
        def("WGB_PRIM_SIZE", prim_size());
        def("WGB_STRIP", is_strip() ? 1 : 0);
        def("WGB_STRIP_SEPARATED", separated ? 1 : 0);
        def("WGB_FRONT_FACE_CW", front_face == WGB_FRONT_FACE_CW ? 1 : 0);
        def("WGB_CULL", cull_mode);
        def("WGB_RESOLVE", mode);
        def("WGB_DEPTH_COMPARE", test ? depth_compare : WGB_COMPARE_ALWAYS);
        def("WGB_DEPTH_WRITE", (test && depth_write) ? 1 : 0);
        def("WGB_HAS_DEPTH", has_depth_attachment ? 1 : 0);
        def("WGB_NUM_COLOR", (long long)targets.size());
        // tuning knobs for A/B runs: WGB_TUNE="NAME=value,NAME=value" becomes #define lines ahead of the kernels
        if (const char* tune = getenv("WGB_TUNE")) {
            std::string t(tune);
            for (size_t at = 0; at < t.size();) {
                size_t end = t.find(',', at);
                if (end == std::string::npos) end = t.size();
                const std::string kv = t.substr(at, end - at);
                const size_t eq = kv.find('=');
                if (eq != std::string::npos && eq > 0) def(kv.substr(0, eq).c_str(), atoll(kv.c_str() + eq + 1));
                at = end + 1;
            }
        }
        for (size_t b = 0; b < vbs.size(); b++)
            for (const auto& a : vbs[b].attrs) {
                snprintf(line, sizeof(line), "#define WGB_ATTR%u_SLOT %zu\n#define WGB_ATTR%u_STRIDE %lluu\n#define WGB_ATTR%u_OFFSET %lluu\n#define WGB_ATTR%u_INSTANCE %s\n",
                         a.shader_location, b, a.shader_location, (unsigned long long)vbs[b].stride, a.shader_location,
                         (unsigned long long)a.offset, a.shader_location, vbs[b].step_mode == WGB_VERTEX_STEP_MODE_INSTANCE ? "true" : "false");
                s += line;
            }
        // blend state of target t: field 0 = enabled, 1..3 = colour src / dst / op, 4..6 = alpha src / dst / op
        s += "#define WGB_HAVE_BLEND_STATE 1\nstatic __device__ __forceinline__ constexpr int wgb_blend_state(int t, int f) { return ";
        for (size_t t = 0; t < targets.size(); t++) {
            const auto& ts = targets[t];
            const bool on = blends() && ts.has_blend;
            snprintf(line, sizeof(line), "t == %zu ? (f == 0 ? %d : f == 1 ? %u : f == 2 ? %u : f == 3 ? %u : f == 4 ? %u : f == 5 ? %u : %u) : ", t, on ? 1 : 0,
                     ts.blend_color.src_factor, ts.blend_color.dst_factor, ts.blend_color.operation,
                     ts.blend_alpha.src_factor, ts.blend_alpha.dst_factor, ts.blend_alpha.operation);
            s += line;
        }
        s += "0; }\n";
        s += "#include \"wgb_prelude.cuh\"\n";
        s += vs_text;
        s += "\n";
        if (has_fragment) s += fs_text;
        else s += "#define WGB_FS_COLOR_MASK 0\n#define WGB_FS_WRITES_FRAG_DEPTH 0\n#define WGB_FS_MAY_DISCARD 0\n#define WGB_FS_EARLY_DEPTH 0\n"
                  "WGB_DEV constexpr int wgb_fs_interp(int) { return 0; }\n"
                  "WGB_DEV bool wgb_fs_entry(const WgbDraw&, const WgbFragIn&, const u32*, WgbFragOut&) { return true; }\n";
        s += "#include \"wgb_raster.cuh\"\n";
        return s;
    }

    std::shared_ptr<KernelSet> variant(bool has_depth_attachment, bool separated) {
        std::lock_guard<std::mutex> lk(mu);
        const uint32_t key = (has_depth_attachment ? 1u : 0u) | (separated ? 2u : 0u);
        auto it = variants.find(key);
        if (it != variants.end()) return it->second;
        const std::string src = build_source(has_depth_attachment, separated);
        variant_source[key] = src;
        std::shared_ptr<KernelSet> ks;
        {
            std::lock_guard<std::recursive_mutex> dl(device->mu);
            auto c = device->kernel_cache.find(src);
            if (c != device->kernel_cache.end()) ks = c->second;
        }
        if (!ks) {
            const std::vector<char> cubin = nvrtc_compile(src);
            ks = std::make_shared<KernelSet>();
            if (!device->compile_only) {
                load_driver_api();
                device->make_current();
                CUresult r = g_drv.ModuleLoadData(&ks->module, cubin.data());
                if (r != CUDA_SUCCESS) fail(WGB_ERROR_DEVICE, "cuModuleLoadData failed: %s", cu_error(r));
                auto fn = [&](const char* n, CUfunction* f) {
                    CUresult q = g_drv.ModuleGetFunction(f, ks->module, n);
                    if (q != CUDA_SUCCESS) fail(WGB_ERROR_DEVICE, "kernel %s missing: %s", n, cu_error(q));
                };
                ks->ordered = src.find("#define WGB_RESOLVE 7\n") != std::string::npos;
                if (ks->ordered) fn("wgb_big_sort_kernel", &ks->big_sort);
                fn("wgb_begin_kernel", &ks->begin);
                fn("wgb_geometry_kernel", &ks->geometry);
                fn("wgb_vertex_kernel", &ks->vertex);
                fn("wgb_geometry_cached_kernel", &ks->geometry_cached);
                fn("wgb_clip_kernel", &ks->clip);
                fn("wgb_scan_kernel", &ks->scan);
                fn("wgb_fill_kernel", &ks->fill);
                fn("wgb_tile_kernel", &ks->tile);
                fn("wgb_clear_kernel", &ks->clear);
                fn("wgb_strip_map_kernel", &ks->strip_map);
            }
            std::lock_guard<std::recursive_mutex> dl(device->mu);
            device->kernel_cache[src] = ks;
        }
        variants[key] = ks;
        return ks;
    }
};

// ---- recorded commands (render_pass/mod.rs:397-423 RenderPassSubCommand) ----
struct SubCommand {
    enum Kind { SetPipeline, SetBindGroup, SetIndexBuffer, SetVertexBuffer, SetViewport, SetScissor, SetBlendConstant, SetStencilReference, Draw, DrawIndexed } kind;
    Ref<RenderPipeline> pipeline;
    Ref<BindGroup> bind_group;
    Ref<Buffer> buffer;
    uint32_t index = 0, index_format = 0;
    uint64_t offset = 0, size = 0;
    float vp[6] = {0, 0, 0, 0, 0, 1};
    uint32_t sc[4] = {0, 0, 0, 0};
    uint32_t first = 0, count = 0, first_instance = 0, instance_count = 0;
    int32_t base_vertex = 0;
    std::vector<uint32_t> dynamic_offsets;
    float blend_constant[4] = {0, 0, 0, 0};
};
struct PassCommand {
    struct Color { Ref<TextureView> view; uint32_t load_op, store_op; double clear[4]; };
    std::vector<Color> colors;
    bool has_depth = false;
    Ref<TextureView> depth_view;
    bool has_depth_ops = false;
    uint32_t depth_load_op = 0;
    float depth_clear = 0.0f;
    bool has_stencil_ops = false;
    std::vector<SubCommand> sub;
};
// encoder-level copies / clears (command.rs:37-115); executed in order with the render passes
struct CopyCommand {
    enum Kind { BufferToBuffer, BufferToTexture, TextureToBuffer, TextureToTexture, ClearBuffer, ClearTexture } kind;
    Ref<Buffer> src_buffer, dst_buffer;
    Ref<Texture> src_texture, dst_texture;
    uint64_t src_offset = 0, dst_offset = 0, size = 0;
    uint32_t src_x = 0, src_y = 0, src_layer = 0, dst_x = 0, dst_y = 0, dst_layer = 0;
    uint32_t bytes_per_row = 0, width = 0, height = 0;
};
// a recorded command is a render pass or a copy (command.rs:173-196 Command)
struct Command { std::shared_ptr<PassCommand> pass; std::shared_ptr<CopyCommand> copy; };
struct CommandBuffer : Object { Ref<Device> device; std::vector<Command> passes; bool submitted = false; };
struct CommandEncoder : Object { Ref<Device> device; std::vector<Command> passes; bool finished = false; std::mutex mu; };
struct RenderPass : Object {
    Ref<CommandEncoder> encoder;
    std::shared_ptr<PassCommand> cmd;
    bool ended = false;
    std::mutex mu;
    void end() {   // render_pass/mod.rs:309-329 (also called from Drop)
        std::lock_guard<std::mutex> lk(mu);
        if (ended) return;
        ended = true;
        std::lock_guard<std::mutex> el(encoder->mu);
        encoder->passes.push_back(Command{cmd, nullptr});
    }
    ~RenderPass() override { end(); }
};

// ------------------------------------------------------------------------------------------
// pass execution (RenderPassCommand::execute, render_pass/mod.rs:345-394)
// ------------------------------------------------------------------------------------------
struct PassState {   // render_pass/state.rs:58-75
    Ref<RenderPipeline> pipeline;
    Ref<BindGroup> bind_groups[WGB_MAX_GROUPS];
    std::vector<uint32_t> dynamic_offsets[WGB_MAX_GROUPS];
    float blend_constant[4] = {0, 0, 0, 0};
    struct Slice { Ref<Buffer> buffer; uint64_t offset = 0, size = 0; };
    Slice vertex_buffers[WGB_MAX_VERTEX_BUFFERS];
    Slice index_buffer;
    uint32_t index_format = 0;
    float vp[6];
    uint32_t sc[4];
};

void launch_on(Device* dev, cudaStream_t stream, CUfunction f, dim3 grid, dim3 block, WgbDraw* d) {
    void* args[] = {d};
    CUresult r = g_drv.LaunchKernel(f, grid.x, grid.y, grid.z, block.x, block.y, block.z, 0, (CUstream)stream, args, nullptr);
    if (r != CUDA_SUCCESS) fail(WGB_ERROR_DEVICE, "kernel launch failed: %s", cu_error(r));
    dev->last_stats.kernel_launches++;
}
void launch(Device* dev, CUfunction f, dim3 grid, dim3 block, WgbDraw* d) { launch_on(dev, dev->joined(), f, grid, block, d); }

void band_rows(const Device* dev, uint32_t tiles_y, uint32_t& ty0, uint32_t& ty1) {
    // contiguous bands of whole tile rows; the tiles_y % n rows left over go to the MIDDLE bands, one each:
    // the first and the last band hold the primitives that cross the top and bottom clip planes (C3 on 8 GPUs: 7 195
    // clipped primitives on the top band against 1 563 on a middle one, 25 us more geometry stage), so they are the
    // ones that can do with a row less
    const uint32_t n = dev->band_count ? dev->band_count : 1u, r = dev->band_rank;
    const uint32_t q = tiles_y / n, rem = tiles_y % n, lo = (n - rem) / 2u;
    auto extra_before = [&](uint32_t k) { return k <= lo ? 0u : (k - lo < rem ? k - lo : rem); };     // taller bands among bands 0 .. k-1
    ty0 = r * q + extra_before(r);
    ty1 = (r + 1u) * q + extra_before(r + 1u);
}

struct PassTargets {
    uint32_t width = 0, height = 0;
    uint32_t num_color = 0;
    WgbAttachment color[WGB_MAX_COLOR];
    bool has_depth = false;
    WgbAttachment depth;
    // tensor maps over colour attachment 0 and the depth attachment (encoded once per pass)
    bool tmap_color = false, tmap_depth = false;
    WgbTensorMap map_color, map_depth;
};

// The attachment (linear, W x H texels of 4 bytes, texture.rs:250-302) as a 2-D tensor map with a tile-sized box.  TMA
// wants a 16-byte aligned base and row pitch; attachments that do not qualify keep the row-by-row bulk copies.
bool encode_attachment_map(const WgbAttachment& a, uint32_t width, uint32_t height, WgbTensorMap& out) {
    static_assert(sizeof(WgbTensorMap) == sizeof(CUtensorMap), "WgbTensorMap mirrors CUtensorMap");
    if (!g_drv.TensorMapEncodeTiled || a.bytes_per_texel != 4 || (a.ptr & 15ull) || (((uint64_t)width * 4) & 15ull) || !width || !height) return false;
    const cuuint64_t dims[2] = {width, height};
    const cuuint64_t strides[1] = {(cuuint64_t)width * 4};
    const cuuint32_t box[2] = {WGB_TILE_W, WGB_TILE_H};
    const cuuint32_t elem[2] = {1, 1};
    return g_drv.TensorMapEncodeTiled(reinterpret_cast<CUtensorMap*>(&out), CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, (void*)(uintptr_t)a.ptr, dims, strides, box, elem,
                                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// What a draw batch left behind, once its kernels have completed: timings, the capacity the fullest tile needed, counters.
// Returns true if a work buffer overflowed (the capacities have been raised: the batch has to run again); raises the
// error a geometry-stage status bit stands for.
bool batch_result(Device* dev, const Device::PendingBatch& pb, uint32_t np, uint32_t band_tiles, wgb_pass_stats& stats) {
    const WgbCounters c = *pb.host;
    float g_ms = 0, t_ms = 0;
    cudaEventElapsedTime(&g_ms, pb.ev[0], pb.ev[1]);
    cudaEventElapsedTime(&t_ms, pb.ev[3], pb.ev[2]);
    stats.geometry_ms += g_ms;
    stats.tile_ms += t_ms;
    stats.total_ms += g_ms + t_ms;
    if (dev->bin_cap_hint.size() > 4096) dev->bin_cap_hint.clear();
    if (c.max_tile_pairs) {     // never shrinks: draws of one shape whose fullest tile varies (a moving camera) must not alternate between overflow and replay
        uint32_t& hint = dev->bin_cap_hint[std::make_pair(np, band_tiles)];
        hint = std::max<uint32_t>(hint, (uint32_t)std::min<uint64_t>((uint64_t)c.max_tile_pairs * 5 / 4 + 64, 0xFFFFFF00ull));
    }
    if (c.status & (WGB_STATUS_CLIP_OVERFLOW | WGB_STATUS_BIG_OVERFLOW | WGB_STATUS_BIN_OVERFLOW)) {
        if (c.status & WGB_STATUS_CLIP_OVERFLOW) dev->clip_capacity = std::max<uint32_t>(dev->clip_capacity * 2, c.num_clip_records + 1024);
        if (c.status & WGB_STATUS_BIG_OVERFLOW) dev->big_capacity = std::max<uint32_t>(dev->big_capacity * 2, c.num_big + 1024);
        return true;
    }
    if (c.status & WGB_STATUS_INDEX_OOB) fail(WGB_ERROR_OUT_OF_BOUNDS, "an index lies outside the bound index buffer or overflows with base_vertex (index.rs:45-62)");
    if (c.status & WGB_STATUS_VERTEX_OOB) fail(WGB_ERROR_OUT_OF_BOUNDS, "a vertex attribute fetch lies outside its vertex buffer (vertex.rs:143-154)");
    if (c.status & WGB_STATUS_W_ZERO) fail(WGB_ERROR_VALIDATION, "a clip position has w = 0 (the reference panics: raster.rs:146, primitive.rs:175-177)");
    stats.fragments += c.fragments;
    stats.shaded += c.shaded;
    stats.bin_pairs += c.num_small_pairs;
    stats.big_primitives += c.num_big;
    stats.clipped_primitives += c.num_slow;
    stats.clip_records += c.num_clip_records;
    stats.hiz_culled += c.hiz_culled;
    return false;
}

void execute_draw(Device* dev, PassState& st, PassTargets& tg, const SubCommand& sc) {
    RenderPipeline* pipe = st.pipeline.get();
    if (!pipe) fail(WGB_ERROR_VALIDATION, "No pipeline bound");                       // state.rs:240-243
    if (!pipe->has_fragment) return;     // no fragment state: only the vertex stage would run (state.rs:583-588)
    const bool indexed = sc.kind == SubCommand::DrawIndexed;
    if (indexed && !st.index_buffer.buffer) fail(WGB_ERROR_VALIDATION, "No index buffer bound");   // state.rs:393-397
    bool separated = false;
    if (indexed && pipe->strip_index_format != WGB_INDEX_FORMAT_NONE) {               // state.rs:296-305, 357-359
        REQUIRE(pipe->is_strip(), "strip_index_format set, but not a strip topology");
        REQUIRE(pipe->strip_index_format == st.index_format, "strip_index_format does not match the bound index format");
        separated = true;
    }
    REQUIRE(tg.num_color == pipe->targets.size(), "pipeline has %zu colour targets, pass has %u attachments", pipe->targets.size(), tg.num_color);
    for (uint32_t c = 0; c < tg.num_color; c++)
        REQUIRE(tg.color[c].format == pipe->targets[c].format, "colour target %u format mismatch", c);
    const int ps = pipe->prim_size();
    uint64_t ppi;
    if (!pipe->is_strip()) ppi = sc.count / ps;
    else ppi = sc.count >= (uint32_t)ps ? sc.count - (ps - 1) : 0;
    if (sc.instance_count == 0 || ppi == 0) return;

    std::shared_ptr<KernelSet> ks = pipe->variant(tg.has_depth, separated);
    if (dev->compile_only) fail(WGB_ERROR_DEVICE, "a compile-only device cannot execute draws");

    WgbDraw d;
    memset(&d, 0, sizeof(d));
    d.fb_width = tg.width; d.fb_height = tg.height;
    d.tiles_x = (tg.width + WGB_TILE_W - 1) / WGB_TILE_W;
    d.tiles_y = (tg.height + WGB_TILE_H - 1) / WGB_TILE_H;
    band_rows(dev, d.tiles_y, d.band_ty0, d.band_ty1);
    // ToRaster::new (raster.rs:129-143)
    d.vp_tx = st.vp[0] + 0.5f * st.vp[2];
    d.vp_ty = st.vp[1] + 0.5f * st.vp[3];
    d.vp_sx = 0.5f * st.vp[2];
    d.vp_sy = 0.5f * st.vp[3];
    d.vp_sy *= -1.0f;
    d.sc_x0 = st.sc[0]; d.sc_y0 = st.sc[1]; d.sc_x1 = st.sc[0] + st.sc[2]; d.sc_y1 = st.sc[1] + st.sc[3];
    d.indexed = indexed ? st.index_format : 0u;
    d.first = sc.first; d.count = sc.count; d.base_vertex = sc.base_vertex;
    d.first_instance = sc.first_instance; d.instance_count = sc.instance_count;
    if (indexed) {
        Buffer* ib = st.index_buffer.buffer.get();
        REQUIRE(ib->device.get() == dev, "index buffer belongs to another device");
        REQUIRE(st.index_buffer.offset <= ib->size, "index buffer offset out of range");
        // bytemuck::cast_slice panics on a misaligned slice (index.rs:45-51); a misaligned device load would be a sticky fault
        REQUIRE(st.index_buffer.offset % (st.index_format == WGB_INDEX_FORMAT_UINT16 ? 2 : 4) == 0, "index buffer offset is not a multiple of the index size");
        d.index_ptr = (uint64_t)(uintptr_t)ib->dptr + st.index_buffer.offset;
        d.index_size = std::min<uint64_t>(st.index_buffer.size, ib->size - st.index_buffer.offset);
        // every index of the range is fetched (state.rs:521-535), so a range that leaves the buffer is a certain
        // slice panic in the reference (index.rs:45-51); checked once here instead of per index on the device
        const uint64_t isz = st.index_format == WGB_INDEX_FORMAT_UINT16 ? 2 : 4;
        if (((uint64_t)sc.first + sc.count) * isz > d.index_size)
            fail(WGB_ERROR_OUT_OF_BOUNDS, "draw_indexed range [%u, %u) lies outside the bound index buffer (%llu bytes)", sc.first,
                 sc.first + sc.count, (unsigned long long)d.index_size);
    }
    for (size_t b = 0; b < pipe->vbs.size(); b++) {                                     // vertex.rs:262-272
        Buffer* vb = st.vertex_buffers[b].buffer.get();
        if (!vb) fail(WGB_ERROR_VALIDATION, "Buffer %zu not bound", b);
        REQUIRE(vb->device.get() == dev, "vertex buffer %zu belongs to another device", b);
        REQUIRE(st.vertex_buffers[b].offset <= vb->size, "vertex buffer offset out of range");
        d.vb[b].ptr = (uint64_t)(uintptr_t)vb->dptr + st.vertex_buffers[b].offset;
        d.vb[b].size = std::min<uint64_t>(st.vertex_buffers[b].size, vb->size - st.vertex_buffers[b].offset);
    }
    Buffer* res_buffer[WGB_MAX_GROUPS][WGB_MAX_BINDINGS] = {};
    for (uint32_t g = 0; g < WGB_MAX_GROUPS; g++) {                                     // binding.rs:22-52
        BindGroup* bg = st.bind_groups[g].get();
        if (!bg) continue;
        // WGB_FEATURE_DYNAMIC_OFFSETS: the k-th offset belongs to the k-th dynamic binding in increasing binding number
        std::map<uint32_t, uint64_t> dyn;
        if ((dev->features & WGB_FEATURE_DYNAMIC_OFFSETS) && bg->layout) {
            std::vector<uint32_t> dynamic_bindings;
            for (const auto& le : bg->layout->entries) if (le.has_dynamic_offset) dynamic_bindings.push_back(le.binding);
            std::sort(dynamic_bindings.begin(), dynamic_bindings.end());
            REQUIRE(dynamic_bindings.size() == st.dynamic_offsets[g].size(), "bind group %u takes %zu dynamic offsets, %zu were given",
                    g, dynamic_bindings.size(), st.dynamic_offsets[g].size());
            for (size_t k = 0; k < dynamic_bindings.size(); k++) dyn[dynamic_bindings[k]] = st.dynamic_offsets[g][k];
        }
        for (const auto& e : bg->entries) {
            if (e.binding >= WGB_MAX_BINDINGS) fail(WGB_ERROR_UNSUPPORTED, "binding index %u exceeds the supported %d", e.binding, WGB_MAX_BINDINGS);
            WgbResource& r = d.res[g][e.binding];
            r.kind = e.kind;
            if (e.kind == WGB_BINDING_BUFFER) {
                REQUIRE(e.buffer->device.get() == dev, "bind group buffer belongs to another device");
                const uint64_t off = e.offset + (dyn.count(e.binding) ? dyn[e.binding] : 0);
                REQUIRE(off <= e.buffer->size, "bind group buffer offset out of range");
                r.ptr = (uint64_t)(uintptr_t)e.buffer->dptr + off;
                r.a = (uint32_t)std::min<uint64_t>(e.size, e.buffer->size - off);
                res_buffer[g][e.binding] = e.buffer.get();
            } else if (e.kind == WGB_BINDING_TEXTURE_VIEW) {
                Texture* t = e.view->texture.get();
                if (t->desc.format != WGB_TEXTURE_FORMAT_RGBA8_UNORM && t->desc.format != WGB_TEXTURE_FORMAT_RGBA8_UNORM_SRGB)
                    fail(WGB_ERROR_UNSUPPORTED, "sampled texture format %u (the reference reads Rgba8Unorm[Srgb] only, texture.rs:176-186)", t->desc.format);
                REQUIRE(t->device.get() == dev, "sampled texture belongs to another device");
                if (e.view->base_layer != 0) fail(WGB_ERROR_UNSUPPORTED, "sampled texture views must start at array layer 0 (the reference samples layer 0 only, texture.rs:176-186)");
                r.ptr = (uint64_t)(uintptr_t)t->dptr;
                r.tex = (uint64_t)t->texture_object();
                r.a = t->desc.width; r.b = t->desc.height; r.c = t->desc.format;
            } else if (e.kind == WGB_BINDING_SAMPLER) {
                r.a = e.sampler->desc.address_mode_u; r.b = e.sampler->desc.address_mode_v;
            }
        }
    }
    // post-transform vertex cache: indexed draws whose per-vertex buffers bound the vertex range
    uint64_t vcache_n = 0;
    if (indexed && !getenv("WGB_NO_VERTEX_CACHE")) {
        uint64_t vn = UINT64_MAX;
        bool any = false;
        for (size_t b = 0; b < pipe->vbs.size(); b++) {
            if (pipe->vbs[b].step_mode != WGB_VERTEX_STEP_MODE_VERTEX || pipe->vbs[b].attrs.empty()) continue;
            uint64_t max_end = 0;
            for (const auto& a : pipe->vbs[b].attrs) max_end = std::max<uint64_t>(max_end, a.offset + vertex_format_size(a.format));
            any = true;
            if (d.vb[b].size < max_end) { vn = 0; break; }
            if (pipe->vbs[b].stride == 0) continue;                  // every vertex reads the same bytes
            vn = std::min<uint64_t>(vn, (d.vb[b].size - max_end) / pipe->vbs[b].stride + 1);
        }
        // worth it only if the cache is not much larger than the index range, and addressable
        if (any && vn != UINT64_MAX && vn > 0 && vn <= 2ull * sc.count + 1024 && vn * sc.instance_count <= (1ull << 27)) vcache_n = vn;
    }
    d.num_color = tg.num_color;
    d.has_depth = tg.has_depth ? 1u : 0u;
    for (uint32_t c = 0; c < tg.num_color; c++) {
        d.color[c] = tg.color[c];
        if (dev->features & WGB_FEATURE_COLOR_WRITE_MASK) d.color[c].write_mask = texel_write_mask(pipe->targets[c].write_mask, tg.color[c].format);
    }
    if (tg.has_depth) d.depth = tg.depth;
    if (tg.tmap_color) { d.tmap_color = 1u; d.map_color = tg.map_color; }
    if (tg.tmap_depth) { d.tmap_depth = 1u; d.map_depth = tg.map_depth; }
    memcpy(d.blend_constant, st.blend_constant, sizeof(d.blend_constant));
    if (dev->features & WGB_FEATURE_VIEWPORT_DEPTH_RANGE) {
        d.depth_range = 1u; d.depth_min = st.vp[4]; d.depth_scale = st.vp[5] - st.vp[4];
    }
    d.stats = dev->coverage_capture ? 1u : 0u;

    const uint32_t band_tiles = d.tiles_x * (d.band_ty1 - d.band_ty0);
    if (band_tiles == 0) return;

    // primitive restart: positions of every primitive's vertices in the index range
    if (separated) {
        dev->strip_map.ensure((size_t)ppi * ps * 4);
        dev->strip_count.ensure(4);
        d.strip_map = dev->strip_map.addr();
        CUDA_CHECK(cudaMemsetAsync(dev->strip_map.p, 0xFF, (size_t)ppi * ps * 4, dev->joined()));
        void* args[] = {&d, &dev->strip_count.p};
        CUresult r = g_drv.LaunchKernel(ks->strip_map, 1, 1, 1, 32, 1, 1, 0, (CUstream)dev->joined(), args, nullptr);
        if (r != CUDA_SUCCESS) fail(WGB_ERROR_DEVICE, "strip map launch failed: %s", cu_error(r));
        dev->last_stats.kernel_launches++;
        uint32_t n = 0;
        CUDA_CHECK(cudaMemcpyAsync(&n, dev->strip_count.p, 4, cudaMemcpyDeviceToHost, dev->joined()));
        CUDA_CHECK(cudaStreamSynchronize(dev->joined()));
        ppi = n;
        if (ppi == 0) return;
    }
    d.prims_per_instance = (uint32_t)ppi;
    const uint64_t total_prims = ppi * (uint64_t)sc.instance_count;
    const uint64_t max_batch = (1ull << 26) - 1;        // order = 1 + 64 * p + sub has to fit 32 bits

    if (dev->coverage_capture && (dev->coverage_w != tg.width || dev->coverage_h != tg.height)) {
        dev->coverage.ensure((size_t)tg.width * tg.height * 4);
        CUDA_CHECK(cudaMemsetAsync(dev->coverage.p, 0, (size_t)tg.width * tg.height * 4, dev->joined()));
        dev->coverage_w = tg.width; dev->coverage_h = tg.height;
    }
    d.coverage = dev->coverage.addr();

    uint64_t live_res[WGB_MAX_GROUPS][WGB_MAX_BINDINGS];          // where the bindings are (a draw that is not waited for reads the small ones from a snapshot)
    for (uint32_t g = 0; g < WGB_MAX_GROUPS; g++) for (uint32_t b = 0; b < WGB_MAX_BINDINGS; b++) live_res[g][b] = d.res[g][b].ptr;
    for (uint64_t base = 0; base < total_prims; base += max_batch) {
        const uint32_t np = (uint32_t)std::min<uint64_t>(max_batch, total_prims - base);
        d.prim_base = (uint32_t)base;   // batches beyond 2^32 primitives are not addressable
        REQUIRE(base + np <= 0xFFFFFFFFull, "draw exceeds 2^32 primitives");
        d.num_prims = np;
        if (dev->clip_capacity == 0) dev->clip_capacity = dev->small_work_buffers ? 2 : 65536;
        for (int attempt = 0;; attempt++) {
            const uint32_t clip_cap = dev->small_work_buffers ? dev->clip_capacity : std::max<uint32_t>(dev->clip_capacity, np / 16);
            const uint32_t big_cap = dev->small_work_buffers ? std::max<uint32_t>(dev->big_capacity, 2) : std::max<uint32_t>(dev->big_capacity, std::max<uint32_t>(65536, np / 8));
            dev->clip_capacity = clip_cap; dev->big_capacity = big_cap;
            // a draw that is not waited for: geometry stage on the render stream without waiting for the tile stream, in
            // the work set the draw before it does not use; its tile kernel follows on the tile stream
            const bool async = dev->recording != nullptr;
            static const bool always_two = getenv("WGB_ALWAYS_TWO_STREAMS") != nullptr;      // A/B knob
            const bool two_streams = async && dev->overlap_stages && (dev->ops_since_tile == 0 || always_two);
            for (uint32_t g = 0; g < WGB_MAX_GROUPS; g++) for (uint32_t b = 0; b < WGB_MAX_BINDINGS; b++) d.res[g][b].ptr = live_res[g][b];
            Device::WorkSet& ws = dev->work[two_streams ? (dev->work_next++ & 1u) : 0u];
            // direct binning: a fixed number of slots per tile, sized from what this draw shape needed before
            // (or twice the mean on first sight); the geometry kernels then store the bin entries themselves and
            // the scan + fill kernels are skipped.  A tile that overflows flags the pass and the draw is replayed.
            uint32_t bin_cap = 0;
            {
                const std::pair<uint32_t, uint32_t> key(np, band_tiles);
                auto it = dev->bin_cap_hint.find(key);
                uint64_t cap = it != dev->bin_cap_hint.end() ? it->second
                                                             : 2 * ((uint64_t)np * 5 / 4) / std::max<uint32_t>(band_tiles, 1) + 64;
                cap = (std::max<uint64_t>(cap, 256) + 255) & ~255ull;
                if (!dev->no_direct_bins && band_tiles > 0 && (uint64_t)band_tiles * cap * 4 <= (1ull << 30)) bin_cap = (uint32_t)cap;
            }
            d.bin_cap = bin_cap;
            // A draw that is not waited for sizes BOTH work sets: the one it does not use is the next draw's, and a first
            // cudaMalloc of 10^8 bytes in the middle of a queue of passes stalls all of them.
            for (Device::WorkSet& w : dev->work) {
                if (&w != &ws && !(async && dev->overlap_stages)) continue;
                w.counters.ensure(sizeof(WgbCounters) + (size_t)(band_tiles + 1) * 4);       // the counters and the per-tile pair counts share one buffer: K0 clears both
                w.prim_box.ensure((size_t)np * 4);
                if (!vcache_n) w.setup_cache.ensure((size_t)np * 48);
                if (vcache_n) {
                    const size_t nv = (size_t)vcache_n * sc.instance_count;
                    w.vcache_raster.ensure(nv * 16); w.vcache_ndc.ensure(nv * 8); w.vcache_flags.ensure(nv);
                }
                w.slow_list.ensure((size_t)np * 4);
                w.clip_records.ensure((size_t)clip_cap * sizeof(WgbClipRecord));
                w.big_list.ensure((size_t)big_cap * sizeof(WgbBigEntry));
                w.tile_offset.ensure((size_t)(band_tiles + 1) * 4);
                w.tile_cursor.ensure((size_t)(band_tiles + 1) * 4);
                w.bins.ensure(bin_cap ? (size_t)band_tiles * bin_cap * 4 : ((size_t)np + clip_cap) * WGB_SMALL_MAX_TILES * 4);
                if (async) w.snapshots.ensure((size_t)WGB_MAX_GROUPS * WGB_MAX_BINDINGS * 1024);
            }
            if (vcache_n) {
                d.vcache_raster = ws.vcache_raster.addr(); d.vcache_ndc = ws.vcache_ndc.addr(); d.vcache_flags = ws.vcache_flags.addr();
                d.vcache_count = (uint32_t)vcache_n;
            }
            d.counters = ws.counters.addr(); d.prim_box = ws.prim_box.addr(); d.slow_list = ws.slow_list.addr();
            d.setup_cache = ws.setup_cache.addr();
            d.clip_records = ws.clip_records.addr(); d.clip_capacity = clip_cap;
            d.big_list = ws.big_list.addr(); d.big_capacity = big_cap;
            d.tile_count = ws.counters.addr() + sizeof(WgbCounters); d.tile_offset = ws.tile_offset.addr(); d.tile_cursor = ws.tile_cursor.addr();
            d.bins = ws.bins.addr();

            d.poison = async ? dev->poison.addr() : 0;
            Device::PendingBatch pb{};
            if (async) {
                pb.host = dev->counter_ring + dev->counter_ring_used++;
                for (auto& e : pb.ev) e = dev->pool_event();
                pb.np = np; pb.band_tiles = band_tiles;
            } else {
                pb.host = dev->host_counters;
                for (int k = 0; k < 3; k++) pb.ev[k] = dev->ev[k];
            }
            cudaStream_t gs, ts;
            if (two_streams) {
                gs = dev->stream;
                ts = dev->tile_stream;
                if (ws.busy) { CUDA_CHECK(cudaStreamWaitEvent(gs, ws.ev_free, 0)); ws.busy = false; }
            } else gs = ts = dev->joined();
            CUDA_CHECK(cudaEventRecord(pb.ev[0], gs));
            // K0: counters and per-tile pair counts to zero (they share one buffer), snapshots of the small bindings
            d.begin_words = (uint32_t)((sizeof(WgbCounters) + (size_t)(band_tiles + 1) * 4) / 4);
            memset(d.snap_src, 0, sizeof(d.snap_src));
            if (async) {
                dev->draw_serial++;
                constexpr uint32_t SNAP_MAX = 1024, SNAP_SLOT = 1024;
                for (uint32_t g = 0; g < WGB_MAX_GROUPS; g++)
                    for (uint32_t b = 0; b < WGB_MAX_BINDINGS; b++) {
                        Buffer* rb = res_buffer[g][b];
                        if (!rb) continue;
                        WgbResource& r = d.res[g][b];
                        static const bool no_snapshots = getenv("WGB_NO_SNAPSHOTS") != nullptr;      // A/B knob
                        if (r.a > 0 && r.a <= SNAP_MAX && !no_snapshots) {
                            d.snap_src[g][b] = live_res[g][b];
                            r.ptr = ws.snapshots.addr() + (uint64_t)(g * WGB_MAX_BINDINGS + b) * SNAP_SLOT;
                        } else rb->last_live_draw = dev->draw_serial;
                    }
                for (size_t b = 0; b < pipe->vbs.size(); b++) st.vertex_buffers[b].buffer->last_live_draw = dev->draw_serial;
                if (indexed) st.index_buffer.buffer->last_live_draw = dev->draw_serial;
            }
            static const bool memset_begin = getenv("WGB_MEMSET_BEGIN") != nullptr;                   // A/B knob (needs WGB_NO_SNAPSHOTS)
            if (memset_begin) CUDA_CHECK(cudaMemsetAsync(ws.counters.p, 0, (size_t)d.begin_words * 4, gs));
            else launch_on(dev, gs, ks->begin, dim3(std::min<uint32_t>((d.begin_words + 255) / 256, 148 * 4)), dim3(256), &d);
            const uint32_t gblocks = (np + 255) / 256;
            if (vcache_n) {
                // the cache holds every instance of the draw, so it is filled once, by the first batch
                if (base == 0) launch_on(dev, gs, ks->vertex, dim3((uint32_t)(((uint64_t)vcache_n * sc.instance_count + 255) / 256)), dim3(256), &d);
                launch_on(dev, gs, ks->geometry_cached, dim3((np + 255) / 256), dim3(256), &d);
            } else launch_on(dev, gs, ks->geometry, dim3(gblocks), dim3(256), &d);
            launch_on(dev, gs, ks->clip, dim3(std::min<uint32_t>(std::max<uint32_t>((np + 127) / 128, 148), 148 * 8)), dim3(128), &d);      // (the kernel deals its list out over the warps it finds)
            if (!bin_cap) {
                launch_on(dev, gs, ks->scan, dim3(1), dim3(1024), &d);
                launch_on(dev, gs, ks->fill, dim3(gblocks), dim3(256), &d);
            }
            if (ks->ordered) launch_on(dev, gs, ks->big_sort, dim3(1), dim3(1024), &d);
            CUDA_CHECK(cudaEventRecord(pb.ev[1], gs));
            if (two_streams) {
                CUDA_CHECK(cudaStreamWaitEvent(ts, pb.ev[1], 0));
                CUDA_CHECK(cudaEventRecord(pb.ev[3], ts));
            } else pb.ev[3] = pb.ev[1];
            launch_on(dev, ts, ks->tile, dim3(d.tiles_x, d.band_ty1 - d.band_ty0), dim3(256), &d);
            CUDA_CHECK(cudaEventRecord(pb.ev[2], ts));
            CUDA_CHECK(cudaMemcpyAsync(pb.host, ws.counters.p, sizeof(WgbCounters), cudaMemcpyDeviceToHost, ts));
            if (two_streams) {
                if (!ws.ev_free) CUDA_CHECK(cudaEventCreateWithFlags(&ws.ev_free, cudaEventDisableTiming));
                CUDA_CHECK(cudaEventRecord(ws.ev_free, ts));
                ws.busy = true;
                dev->tiles_ahead = true;
            }
            dev->ops_since_tile = 0;
            if (async) {
                // not waited for: settle() looks at the counters when somebody waits, and re-runs the pass if this batch failed
                dev->recording->batches.push_back(pb);
                dev->last_stats.primitives += np;
                break;
            }
            CUDA_CHECK(cudaStreamSynchronize(dev->joined()));
            if (batch_result(dev, pb, np, band_tiles, dev->last_stats)) {
                // the tile kernel saw the overflow flag and left the attachments untouched: the buffers have grown, replay
                if (attempt >= 8) fail(WGB_ERROR_OUT_OF_MEMORY, "work buffers still too small after %d replays", attempt);
                dev->last_stats.replays++;
                continue;
            }
            dev->last_stats.primitives += np;
            break;
        }
        // the clear has been applied by the first executed batch
        for (uint32_t c = 0; c < tg.num_color; c++) { tg.color[c].load_clear = 0; d.color[c].load_clear = 0; }
        if (tg.has_depth) { tg.depth.load_clear = 0; d.depth.load_clear = 0; }
    }
    dev->last_stats.draws++;
}

// Tracing: the reference logs `render pass time` and its per-draw counters at debug level through the
// `tracing` crate (render_pass/mod.rs:346,392-393, state.rs:516-517,592; RUST_LOG selects the level).  Here WGB_LOG=debug
// prints one line per executed pass to stderr, WGB_LOG=trace one more per draw, and every pass / draw is an NVTX range
// for nsys / ncu when the toolkit's header is there.
int log_level() {
    static const int level = [] {
        const char* e = getenv("WGB_LOG");
        return !e ? 0 : !strcmp(e, "trace") ? 2 : (!strcmp(e, "debug") || !strcmp(e, "1")) ? 1 : 0;
    }();
    return level;
}
struct TraceRange {
#ifdef WGB_HAVE_NVTX
    explicit TraceRange(const char* name) { nvtxRangePushA(name); }
    ~TraceRange() { nvtxRangePop(); }
#else
    explicit TraceRange(const char*) {}
#endif
};

void execute_pass_body(Device* dev, const PassCommand& pass);

// may this pass be enqueued without waiting for its draws?  It has to be safe to run again from its start: every
// attachment is cleared on load, so a second run starts from the same texels whatever the first one wrote.
bool pass_can_run_async(const Device* dev, const PassCommand& pass) {
    if (!dev->async_submit || dev->compile_only || dev->coverage_capture) return false;
    if (pass.colors.empty() && !pass.has_depth) return false;
    for (const auto& c : pass.colors) if (c.load_op != WGB_LOAD_OP_CLEAR) return false;
    if (pass.has_depth && !(pass.has_depth_ops && pass.depth_load_op == WGB_LOAD_OP_CLEAR)) return false;
    size_t draws = 0;
    for (const SubCommand& sc : pass.sub) {
        if (sc.kind == SubCommand::SetPipeline && sc.pipeline && sc.pipeline->strip_index_format != WGB_INDEX_FORMAT_NONE) return false;   // needs the strip map's count on the host
        if (sc.kind == SubCommand::Draw || sc.kind == SubCommand::DrawIndexed) {
            // (a draw of more than 2^26 primitives is split into batches; keep those synchronous)
            if ((uint64_t)sc.count * std::max<uint32_t>(sc.instance_count, 1u) > (1ull << 26)) return false;
            draws++;
        }
    }
    return draws + 8 < Device::COUNTER_RING;
}

void execute_pass(Device* dev, const std::shared_ptr<PassCommand>& pass, bool allow_async);

// Wait for the asynchronous passes and look at what they left behind (see Device::unsettled).
void settle(Device* dev) {
    if (dev->unsettled.empty() || dev->settling) return;
    dev->settling = true;
    struct Done { Device* d; ~Done() { d->settling = false; d->counter_ring_used = 0; d->event_pool_used = 0; } } done{dev};
    CUDA_CHECK(cudaStreamSynchronize(dev->joined()));
    std::deque<Device::Unsettled> work;
    work.swap(dev->unsettled);
    auto defer = [&](const Error& e) {
        if (dev->deferred_status == WGB_OK) { dev->deferred_status = e.status; dev->deferred_error = e.what(); }
    };
    // whatever happens below, the queued writes are what their regions hold when this returns
    struct Fold { Device* d; std::deque<Device::Unsettled>& w; ~Fold() { for (auto& u : w) if (u.write_buffer) d->note_base_write(u.write_buffer, u.write_offset, u.write_data.data(), u.write_data.size()); } } fold{dev, work};
    size_t i = 0;
    bool overflow = false, error = false;
    for (; i < work.size(); i++) {
        Device::Unsettled& u = work[i];
        if (!u.pass) continue;
        try {
            for (const auto& b : u.batches)
                if (batch_result(dev, b, b.np, b.band_tiles, u.stats)) { overflow = true; break; }
        } catch (const Error& e) { defer(e); error = true; }
        if (overflow || error) break;
        dev->last_stats = u.stats;
    }
    if (i == work.size()) return;
    // Pass i failed; its tile kernel raised the poison word, so no later pass has touched an attachment.  Bring the
    // small writes that were queued since the last settle back into their order, then run pass i (unless it failed
    // with an error: then the rest of its submission is dropped, as when the error is raised synchronously) and
    // everything behind it again, waiting for every draw.
    CUDA_CHECK(cudaMemsetAsync(dev->poison.p, 0, 4, dev->joined()));
    auto rewrite = [&](const Device::Unsettled& u) {
        CUDA_CHECK(cudaMemcpyAsync((char*)u.write_buffer->dptr + u.write_offset, u.write_data.data(), u.write_data.size(), cudaMemcpyHostToDevice, dev->joined()));
    };
    // (first the regions as they were at the head of the log: writes queued behind pass i have gone over them)
    for (const auto& u : work)
        if (u.write_buffer)
            if (const Device::BaseWrite* w = dev->find_base_write(u.write_buffer, u.write_offset, u.write_data.size()))
                CUDA_CHECK(cudaMemcpyAsync((char*)w->buffer->dptr + w->offset, w->data.data(), w->data.size(), cudaMemcpyHostToDevice, dev->joined()));
    for (size_t k = 0; k < i; k++) if (work[k].write_buffer) rewrite(work[k]);
    size_t j = i;
    uint64_t dropped = 0;
    if (error) { dropped = work[i].submission; j = i + 1; }
    for (; j < work.size(); j++) {
        Device::Unsettled& u = work[j];
        if (u.write_buffer) { rewrite(u); continue; }
        if (u.submission == dropped) continue;
        try {
            execute_pass(dev, u.pass, false);
            if (j == i) dev->last_stats.replays++;          // the attempt that overflowed
        } catch (const Error& e) { defer(e); dropped = u.submission; }
    }
    CUDA_CHECK(cudaStreamSynchronize(dev->joined()));
}

void execute_pass(Device* dev, const std::shared_ptr<PassCommand>& pass_ptr, bool allow_async) {
    const PassCommand& pass = *pass_ptr;
    TraceRange range("wgb::render_pass");
    const auto t0 = std::chrono::steady_clock::now();
    const bool async = allow_async && !dev->settling && pass_can_run_async(dev, pass);
    if (async) {
        size_t draws = 0;
        for (const SubCommand& sc : pass.sub) draws += (sc.kind == SubCommand::Draw || sc.kind == SubCommand::DrawIndexed) ? 1 : 0;
        if (dev->counter_ring_used + draws + 8 > Device::COUNTER_RING) settle(dev);
        if (!dev->counter_ring) CUDA_CHECK(cudaMallocHost((void**)&dev->counter_ring, sizeof(WgbCounters) * Device::COUNTER_RING));
        if (!dev->poison.p) { dev->poison.ensure(4); CUDA_CHECK(cudaMemsetAsync(dev->poison.p, 0, 4, dev->joined())); }
        dev->unsettled.emplace_back();
        dev->unsettled.back().submission = dev->current_submission;
        dev->unsettled.back().pass = pass_ptr;
        dev->recording = &dev->unsettled.back();
    } else if (!dev->settling) {
        settle(dev);            // a pass that waits for its draws runs after everything queued before it has been looked at
    }
    struct Stop { Device* d; ~Stop() { if (d->recording) { d->recording->stats = d->last_stats; d->recording = nullptr; } } } stop{dev};
    execute_pass_body(dev, pass);
    if (log_level() >= 1) {
        const wgb_pass_stats& s = dev->last_stats;
        const double host_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (async)
            fprintf(stderr, "wgpu-b200 DEBUG render pass enqueued: %.3f ms on the host; draws=%u primitives_drawn=%llu launches=%u (counters at the next wait)\n",
                    host_ms, s.draws, (unsigned long long)s.primitives, s.kernel_launches);
        else
        fprintf(stderr, "wgpu-b200 DEBUG render pass time: %.3f ms on the device (geometry %.3f, tile %.3f), %.3f ms on the host; draws=%u "
                        "primitives_drawn=%llu fragments=%llu shaded=%llu bin_pairs=%llu big=%llu clipped=%llu hiz_culled=%llu launches=%u replays=%u\n",
                s.total_ms, s.geometry_ms, s.tile_ms, host_ms, s.draws, (unsigned long long)s.primitives, (unsigned long long)s.fragments,
                (unsigned long long)s.shaded, (unsigned long long)s.bin_pairs, (unsigned long long)s.big_primitives,
                (unsigned long long)s.clipped_primitives, (unsigned long long)s.hiz_culled, s.kernel_launches, s.replays);
    }
}
void execute_pass_body(Device* dev, const PassCommand& pass) {
    dev->last_stats = wgb_pass_stats{};
    PassTargets tg;
    bool have_size = false;
    auto check_size = [&](const Texture* t) {                                          // state.rs:84-96
        if (have_size) REQUIRE(t->desc.width == tg.width && t->desc.height == tg.height, "All render attachments must be the same size");
        tg.width = t->desc.width; tg.height = t->desc.height; have_size = true;
    };
    for (const auto& c : pass.colors) {
        if (!c.view) fail(WGB_ERROR_UNSUPPORTED, "empty colour attachment slots are not supported");
        Texture* t = c.view->texture.get();
        REQUIRE(is_color_format(t->desc.format), "colour attachment has a non-colour format");
        REQUIRE(t->device.get() == dev, "attachment belongs to another device");
        check_size(t);
        REQUIRE(tg.num_color < WGB_MAX_COLOR, "too many colour attachments");
        WgbAttachment& a = tg.color[tg.num_color++];
        t->acquire_for_write(dev->stream);      // (a wait only: the render stream is not joined with the tile stream for it)
        a.ptr = (uint64_t)(uintptr_t)t->dptr + (uint64_t)c.view->base_layer * t->desc.width * t->desc.height * t->bpp;
        a.format = t->desc.format;
        a.bytes_per_texel = t->bpp;
        a.load_clear = c.load_op == WGB_LOAD_OP_CLEAR ? 1u : 0u;
        a.srgb_encode = ((dev->features & WGB_FEATURE_SRGB_ENCODE) && is_srgb_format(t->desc.format)) ? 1u : 0u;
        a.write_mask = 0xFFFFFFFFu;
        a.clear_texel = encode_color(c.clear, t->desc.format, a.srgb_encode != 0);
    }
    if (pass.has_depth) {
        Texture* t = pass.depth_view->texture.get();
        if (pass.has_stencil_ops) fail(WGB_ERROR_UNSUPPORTED, "stencil_ops (fragment.rs:618-620 todo!)");
        if (t->desc.format != WGB_TEXTURE_FORMAT_DEPTH32_FLOAT) fail(WGB_ERROR_UNSUPPORTED, "depth attachments must be Depth32Float (texture.rs:220-239 reads raw f32)");
        REQUIRE(t->device.get() == dev, "depth attachment belongs to another device");
        check_size(t);
        tg.has_depth = true;
        t->acquire_for_write(dev->stream);
        tg.depth.ptr = (uint64_t)(uintptr_t)t->dptr + (uint64_t)pass.depth_view->base_layer * t->desc.width * t->desc.height * 4ull;
        tg.depth.format = t->desc.format;
        tg.depth.bytes_per_texel = 4;
        tg.depth.load_clear = (pass.has_depth_ops && pass.depth_load_op == WGB_LOAD_OP_CLEAR) ? 1u : 0u;
        memcpy(&tg.depth.clear_texel, &pass.depth_clear, 4);
    }
    // (a target mapped from another process keeps the row copies: peer memory is written with plain bulk stores)
    if (!dev->compile_only && tg.num_color >= 1 && !pass.colors[0].view->texture->imported)
        tg.tmap_color = encode_attachment_map(tg.color[0], tg.width, tg.height, tg.map_color);
    if (!dev->compile_only && tg.has_depth) tg.tmap_depth = encode_attachment_map(tg.depth, tg.width, tg.height, tg.map_depth);
    PassState st;
    // default viewport / scissor: the whole framebuffer (state.rs:604-628)
    st.vp[0] = 0; st.vp[1] = 0; st.vp[2] = (float)tg.width; st.vp[3] = (float)tg.height; st.vp[4] = 0; st.vp[5] = 1;
    st.sc[0] = 0; st.sc[1] = 0; st.sc[2] = tg.width; st.sc[3] = tg.height;

    if (dev->coverage_capture && have_size) {
        dev->coverage.ensure((size_t)tg.width * tg.height * 4);
        CUDA_CHECK(cudaMemsetAsync(dev->coverage.p, 0, (size_t)tg.width * tg.height * 4, dev->joined()));
        dev->coverage_w = tg.width; dev->coverage_h = tg.height;
    }

    for (const SubCommand& sc : pass.sub) {                                            // render_pass/mod.rs:351-388
        switch (sc.kind) {
            case SubCommand::SetPipeline: st.pipeline = sc.pipeline; break;
            case SubCommand::SetBindGroup:
                if (sc.index >= WGB_MAX_GROUPS) fail(WGB_ERROR_UNSUPPORTED, "bind group index %u", sc.index);
                st.bind_groups[sc.index] = sc.bind_group;          // dynamic offsets are stored then ignored (state.rs:194-205)
                st.dynamic_offsets[sc.index] = sc.dynamic_offsets;  // ... unless WGB_FEATURE_DYNAMIC_OFFSETS was requested
                break;
            case SubCommand::SetIndexBuffer:
                st.index_buffer.buffer = sc.buffer; st.index_buffer.offset = sc.offset; st.index_buffer.size = sc.size;
                st.index_format = sc.index_format;
                break;
            case SubCommand::SetVertexBuffer:
                if (sc.index >= WGB_MAX_VERTEX_BUFFERS) fail(WGB_ERROR_UNSUPPORTED, "vertex buffer slot %u", sc.index);
                st.vertex_buffers[sc.index].buffer = sc.buffer; st.vertex_buffers[sc.index].offset = sc.offset; st.vertex_buffers[sc.index].size = sc.size;
                break;
            case SubCommand::SetViewport: memcpy(st.vp, sc.vp, sizeof(st.vp)); break;
            case SubCommand::SetScissor: memcpy(st.sc, sc.sc, sizeof(st.sc)); break;
            case SubCommand::SetBlendConstant: memcpy(st.blend_constant, sc.blend_constant, sizeof(st.blend_constant)); break;   // used by WGB_FEATURE_BLEND only
            case SubCommand::SetStencilReference: break;   // stored, unused (state.rs:207-221)
            case SubCommand::Draw: case SubCommand::DrawIndexed: {
                TraceRange draw_range(sc.kind == SubCommand::Draw ? "wgb::draw" : "wgb::draw_indexed");
                const wgb_pass_stats before = dev->last_stats;
                execute_draw(dev, st, tg, sc);
                if (log_level() >= 2)
                    fprintf(stderr, "wgpu-b200 TRACE %s: primitives=%llu fragments=%llu geometry %.3f ms tile %.3f ms replays=%u\n",
                            sc.kind == SubCommand::Draw ? "draw" : "draw_indexed", (unsigned long long)(dev->last_stats.primitives - before.primitives),
                            (unsigned long long)(dev->last_stats.fragments - before.fragments), dev->last_stats.geometry_ms - before.geometry_ms,
                            dev->last_stats.tile_ms - before.tile_ms, dev->last_stats.replays - before.replays);
                break;
            }
        }
    }
    // LoadOp::Clear of a pass whose draws never reached the tile kernel (State::load, state.rs:135-145)
    bool pending = tg.has_depth && tg.depth.load_clear;
    for (uint32_t c = 0; c < tg.num_color; c++) pending = pending || tg.color[c].load_clear;
    if (pending && have_size && !dev->compile_only) {
        // any compiled variant carries the clear kernel; build a minimal one if no pipeline was used
        static const char* kClearSrc =
            "#define WGB_PRIM_SIZE 3\n#define WGB_STRIP 0\n#define WGB_STRIP_SEPARATED 0\n#define WGB_FRONT_FACE_CW 0\n#define WGB_CULL 0\n"
            "#define WGB_RESOLVE 0\n#define WGB_DEPTH_COMPARE 8\n#define WGB_DEPTH_WRITE 0\n#define WGB_HAS_DEPTH 0\n#define WGB_NUM_COLOR 0\n"
            "#include \"wgb_prelude.cuh\"\n#define WGB_VS_VARYING_SLOTS 0\n"
            "WGB_DEV void wgb_vs_entry(const WgbDraw&, u32, u32, vec4f& p, u32*, u32&) { p = vec4f(); }\n"
            "#define WGB_FS_COLOR_MASK 0\n#define WGB_FS_WRITES_FRAG_DEPTH 0\n#define WGB_FS_MAY_DISCARD 0\n#define WGB_FS_EARLY_DEPTH 0\n"
            "WGB_DEV constexpr int wgb_fs_interp(int) { return 0; }\n"
            "WGB_DEV bool wgb_fs_entry(const WgbDraw&, const WgbFragIn&, const u32*, WgbFragOut&) { return true; }\n"
            "#include \"wgb_raster.cuh\"\n";
        std::shared_ptr<KernelSet> ks;
        {
            auto it = dev->kernel_cache.find(kClearSrc);
            if (it != dev->kernel_cache.end()) ks = it->second;
        }
        if (!ks) {
            const std::vector<char> cubin = nvrtc_compile(kClearSrc);
            ks = std::make_shared<KernelSet>();
            load_driver_api();
            CUresult r = g_drv.ModuleLoadData(&ks->module, cubin.data());
            if (r != CUDA_SUCCESS) fail(WGB_ERROR_DEVICE, "cuModuleLoadData failed: %s", cu_error(r));
            r = g_drv.ModuleGetFunction(&ks->clear, ks->module, "wgb_clear_kernel");
            if (r != CUDA_SUCCESS) fail(WGB_ERROR_DEVICE, "clear kernel missing: %s", cu_error(r));
            dev->kernel_cache[kClearSrc] = ks;
        }
        WgbDraw d;
        memset(&d, 0, sizeof(d));
        d.fb_width = tg.width; d.fb_height = tg.height;
        d.tiles_x = (tg.width + WGB_TILE_W - 1) / WGB_TILE_W;
        d.tiles_y = (tg.height + WGB_TILE_H - 1) / WGB_TILE_H;
        band_rows(dev, d.tiles_y, d.band_ty0, d.band_ty1);
        d.num_color = tg.num_color; d.has_depth = tg.has_depth;
        for (uint32_t c = 0; c < tg.num_color; c++) d.color[c] = tg.color[c];
        if (tg.has_depth) d.depth = tg.depth;
        if (dev->recording) {       // asynchronous pass: not waited for (and not timed)
            d.poison = dev->poison.addr();
            launch(dev, ks->clear, dim3(148 * 4), dim3(256), &d);
        } else {
            CUDA_CHECK(cudaEventRecord(dev->ev[0], dev->joined()));
            launch(dev, ks->clear, dim3(148 * 4), dim3(256), &d);
            CUDA_CHECK(cudaEventRecord(dev->ev[2], dev->joined()));
            CUDA_CHECK(cudaStreamSynchronize(dev->joined()));
            float ms = 0;
            cudaEventElapsedTime(&ms, dev->ev[0], dev->ev[2]);
            dev->last_stats.total_ms += ms;
        }
    }
}

void execute_copy(Device* dev, const CopyCommand& c) {
    auto tex_ptr = [](Texture* t, uint32_t x, uint32_t y, uint32_t layer) {
        return (char*)t->dptr + (((uint64_t)layer * t->desc.height + y) * t->desc.width + x) * t->bpp;
    };
    switch (c.kind) {
        case CopyCommand::BufferToBuffer:
            dev->forget_base_writes(c.dst_buffer.get());
            if (c.size) CUDA_CHECK(cudaMemcpyAsync((char*)c.dst_buffer->dptr + c.dst_offset, (char*)c.src_buffer->dptr + c.src_offset, c.size, cudaMemcpyDeviceToDevice, dev->joined()));
            break;
        case CopyCommand::ClearBuffer:
            dev->forget_base_writes(c.dst_buffer.get());
            if (c.size) CUDA_CHECK(cudaMemsetAsync((char*)c.dst_buffer->dptr + c.dst_offset, 0, c.size, dev->joined()));
            break;
        case CopyCommand::ClearTexture:
            c.dst_texture->acquire_for_write(dev->joined());
            CUDA_CHECK(cudaMemsetAsync(c.dst_texture->dptr, 0, c.dst_texture->size, dev->joined()));
            break;
        case CopyCommand::BufferToTexture: {
            Texture* t = c.dst_texture.get();
            t->acquire_for_write(dev->joined());
            const size_t row = (size_t)c.width * t->bpp, pitch = c.bytes_per_row ? c.bytes_per_row : row;
            if (c.width && c.height)
                CUDA_CHECK(cudaMemcpy2DAsync(tex_ptr(t, c.dst_x, c.dst_y, c.dst_layer), (size_t)t->desc.width * t->bpp,
                                             (char*)c.src_buffer->dptr + c.src_offset, pitch, row, c.height, cudaMemcpyDeviceToDevice, dev->joined()));
            break;
        }
        case CopyCommand::TextureToBuffer: {
            Texture* t = c.src_texture.get();
            dev->forget_base_writes(c.dst_buffer.get());
            const size_t row = (size_t)c.width * t->bpp, pitch = c.bytes_per_row ? c.bytes_per_row : row;
            if (c.width && c.height)
                CUDA_CHECK(cudaMemcpy2DAsync((char*)c.dst_buffer->dptr + c.dst_offset, pitch, tex_ptr(t, c.src_x, c.src_y, c.src_layer),
                                             (size_t)t->desc.width * t->bpp, row, c.height, cudaMemcpyDeviceToDevice, dev->joined()));
            break;
        }
        case CopyCommand::TextureToTexture: {
            Texture *st = c.src_texture.get(), *dt = c.dst_texture.get();
            dt->acquire_for_write(dev->joined());
            if (c.width && c.height)
                CUDA_CHECK(cudaMemcpy2DAsync(tex_ptr(dt, c.dst_x, c.dst_y, c.dst_layer), (size_t)dt->desc.width * dt->bpp,
                                             tex_ptr(st, c.src_x, c.src_y, c.src_layer), (size_t)st->desc.width * st->bpp,
                                             (size_t)c.width * st->bpp, c.height, cudaMemcpyDeviceToDevice, dev->joined()));
            break;
        }
    }
}

void record_copy(wgb_command_encoder encoder, std::shared_ptr<CopyCommand> c) {
    CommandEncoder* enc = from_handle<CommandEncoder>(encoder, "command encoder");
    std::lock_guard<std::mutex> lk(enc->mu);
    REQUIRE(!enc->finished, "command encoder is finished");
    enc->passes.push_back(Command{nullptr, c});
}
void check_texture_rect(const Texture* t, uint32_t x, uint32_t y, uint32_t layer, uint32_t w, uint32_t h, const char* what) {
    REQUIRE((uint64_t)x + w <= t->desc.width && (uint64_t)y + h <= t->desc.height && layer < t->desc.depth_or_array_layers,
            "%s: texture rectangle out of bounds", what);
}
void check_buffer_rect(const Buffer* b, uint64_t offset, uint32_t bytes_per_row, uint32_t bpp, uint32_t w, uint32_t h, const char* what) {
    const uint64_t row = (uint64_t)w * bpp, pitch = bytes_per_row ? bytes_per_row : row;
    REQUIRE(pitch >= row, "%s: bytes_per_row smaller than a row", what);
    REQUIRE(h == 0 || range_ok(offset, (uint64_t)(h - 1) * pitch + row, b->size), "%s: buffer range out of bounds", what);
}


// ---- PNG (stored deflate blocks: no compression library in the dependency list) ----
uint32_t crc32_update(uint32_t crc, const uint8_t* p, size_t n) {
    static uint32_t table[256];
    static std::once_flag once;
    std::call_once(once, [] {
        for (uint32_t i = 0; i < 256; i++) {
            uint32_t c = i;
            for (int k = 0; k < 8; k++) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
            table[i] = c;
        }
    });
    for (size_t i = 0; i < n; i++) crc = table[(crc ^ p[i]) & 0xFF] ^ (crc >> 8);
    return crc;
}
void png_chunk(FILE* f, const char type[4], const std::vector<uint8_t>& data) {
    uint8_t len[4] = {(uint8_t)(data.size() >> 24), (uint8_t)(data.size() >> 16), (uint8_t)(data.size() >> 8), (uint8_t)data.size()};
    uint32_t crc = crc32_update(0xFFFFFFFFu, reinterpret_cast<const uint8_t*>(type), 4);
    crc = crc32_update(crc, data.data(), data.size()) ^ 0xFFFFFFFFu;
    uint8_t c[4] = {(uint8_t)(crc >> 24), (uint8_t)(crc >> 16), (uint8_t)(crc >> 8), (uint8_t)crc};
    if (fwrite(len, 1, 4, f) != 4 || fwrite(type, 1, 4, f) != 4 || (data.size() && fwrite(data.data(), 1, data.size(), f) != data.size()) ||
        fwrite(c, 1, 4, f) != 4)
        fail(WGB_ERROR_DEVICE, "short write while saving a PNG");
}
void write_png(const char* path, const uint8_t* px, uint32_t w, uint32_t h, uint32_t channels) {
    REQUIRE(path && px, "null argument");
    REQUIRE(w > 0 && h > 0 && (channels == 1 || channels == 3 || channels == 4), "unsupported PNG layout");
    const size_t row = (size_t)w * channels;
    std::vector<uint8_t> raw;                     // filter byte 0 + pixels, per row
    raw.reserve((row + 1) * h);
    for (uint32_t y = 0; y < h; y++) { raw.push_back(0); raw.insert(raw.end(), px + y * row, px + (y + 1) * row); }
    std::vector<uint8_t> z;                        // zlib stream of stored blocks
    z.reserve(raw.size() + raw.size() / 65535 * 5 + 16);
    z.push_back(0x78); z.push_back(0x01);
    uint32_t a = 1, b = 0;                         // Adler-32
    for (size_t off = 0; off < raw.size();) {
        const size_t n = std::min<size_t>(65535, raw.size() - off);
        z.push_back(off + n == raw.size() ? 1 : 0);
        z.push_back((uint8_t)n); z.push_back((uint8_t)(n >> 8)); z.push_back((uint8_t)~n); z.push_back((uint8_t)(~n >> 8));
        z.insert(z.end(), raw.begin() + off, raw.begin() + off + n);
        for (size_t i = 0; i < n; i += 5552) {
            const size_t m = std::min<size_t>(5552, n - i);
            for (size_t k = 0; k < m; k++) { a += raw[off + i + k]; b += a; }
            a %= 65521; b %= 65521;
        }
        off += n;
    }
    const uint32_t adler = (b << 16) | a;
    z.push_back((uint8_t)(adler >> 24)); z.push_back((uint8_t)(adler >> 16)); z.push_back((uint8_t)(adler >> 8)); z.push_back((uint8_t)adler);
    FILE* f = fopen(path, "wb");
    if (!f) fail(WGB_ERROR_VALIDATION, "cannot open %s for writing", path);
    struct Close { FILE* f; ~Close() { fclose(f); } } closer{f};
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    if (fwrite(sig, 1, 8, f) != 8) fail(WGB_ERROR_DEVICE, "short write while saving a PNG");
    std::vector<uint8_t> ihdr = {(uint8_t)(w >> 24), (uint8_t)(w >> 16), (uint8_t)(w >> 8), (uint8_t)w,
                                 (uint8_t)(h >> 24), (uint8_t)(h >> 16), (uint8_t)(h >> 8), (uint8_t)h,
                                 8, (uint8_t)(channels == 1 ? 0 : channels == 3 ? 2 : 6), 0, 0, 0};
    png_chunk(f, "IHDR", ihdr);
    png_chunk(f, "IDAT", z);
    png_chunk(f, "IEND", {});
}
}  // namespace

// ==========================================================================================
// C ABI
// ==========================================================================================
extern "C" {

const char* wgb_last_error(void) { return g_last_error.c_str(); }
const char* wgb_version(void) { return "wgpu-b200 1 sm_100a"; }
void wgb_retain(wgb_object obj) { if (obj) reinterpret_cast<Object*>(obj)->rc.fetch_add(1); }
void wgb_release(wgb_object obj) {
    if (!obj) return;
    Object* o = reinterpret_cast<Object*>(obj);
    if (o->rc.fetch_sub(1) == 1) delete o;
}
void wgb_free(void* p) { free(p); }

wgb_status wgb_create_instance(const wgb_instance_config*, wgb_instance* out) {
    return guarded([&] { REQUIRE(out, "out is null"); *out = to_handle<wgb_instance>(new Instance()); });
}
wgb_status wgb_instance_request_adapter(wgb_instance instance, wgb_adapter* out) {
    return guarded([&] {
        REQUIRE(out, "out is null");
        Instance* i = from_handle<Instance>(instance, "instance");
        Adapter* a = new Adapter();
        a->instance = Ref<Instance>(i);
        *out = to_handle<wgb_adapter>(a);
    });
}
wgb_status wgb_adapter_get_info(wgb_adapter adapter, wgb_adapter_info* out) {
    return guarded([&] {
        from_handle<Adapter>(adapter, "adapter");
        REQUIRE(out, "out is null");
        memset(out, 0, sizeof(*out));
        int n = 0;
        if (cudaGetDeviceCount(&n) != cudaSuccess) { n = 0; cudaGetLastError(); }
        out->cuda_device_count = (uint32_t)n;
        out->device_type = 1;
        if (n > 0) {
            cudaDeviceProp p;
            int cur = 0;
            cudaGetDevice(&cur);
            if (cudaGetDeviceProperties(&p, cur) == cudaSuccess) snprintf(out->name, sizeof(out->name), "wgpu-b200 (%.100s)", p.name);
        }
        if (!out->name[0]) snprintf(out->name, sizeof(out->name), "wgpu-b200 (no CUDA device)");
    });
}

#define WGB_CUDA_DEVICE_COMPILE_ONLY (-2)
wgb_status wgb_adapter_request_device(wgb_adapter adapter, const wgb_device_descriptor* desc, wgb_device* out_device, wgb_queue* out_queue) {
    return guarded([&] {
        from_handle<Adapter>(adapter, "adapter");
        REQUIRE(out_device && out_queue, "out is null");
        wgb_device_descriptor dd{-1, 0, 1};
        if (desc) dd = *desc;
        Ref<Device> dev;
        dev.p = new Device();
        if (dd.cuda_device == WGB_CUDA_DEVICE_COMPILE_ONLY) {
            dev->compile_only = true;     // shader translation + NVRTC only; cannot allocate or draw
        } else {
            int n = 0;
            cudaError_t e = cudaGetDeviceCount(&n);
            if (e != cudaSuccess || n == 0) { cudaGetLastError(); fail(WGB_ERROR_DEVICE, "no CUDA device available (%s); this backend has no CPU fallback", cudaGetErrorString(e)); }
            int ord = dd.cuda_device;
            if (ord < 0) CUDA_CHECK(cudaGetDevice(&ord));
            REQUIRE(ord < n, "CUDA device %d does not exist (%d devices)", ord, n);
            dev->ordinal = ord;
            CUDA_CHECK(cudaSetDevice(ord));
            CUDA_CHECK(cudaFree(0));
            cudaDeviceProp p;
            CUDA_CHECK(cudaGetDeviceProperties(&p, ord));
            if (p.major != 10) fail(WGB_ERROR_DEVICE, "device %d is sm_%d%d; this backend is built for sm_100a (B200) only", ord, p.major, p.minor);
            load_driver_api();
            {
                int prio_least = 0, prio_greatest = 0;
                CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
                const char* pm = getenv("WGB_STREAM_PRIORITY");      // A/B knob: "flags" (no priorities), "equal", default: render stream first
                if (pm && !strcmp(pm, "flags")) {
                    CUDA_CHECK(cudaStreamCreateWithFlags(&dev->stream, cudaStreamNonBlocking));
                    CUDA_CHECK(cudaStreamCreateWithFlags(&dev->tile_stream, cudaStreamNonBlocking));
                } else {
                    CUDA_CHECK(cudaStreamCreateWithPriority(&dev->stream, cudaStreamNonBlocking, (pm && !strcmp(pm, "equal")) ? prio_least : prio_greatest));
                    CUDA_CHECK(cudaStreamCreateWithPriority(&dev->tile_stream, cudaStreamNonBlocking, prio_least));
                }
            }
            CUDA_CHECK(cudaEventCreateWithFlags(&dev->ev_tiles, cudaEventDisableTiming));
            CUDA_CHECK(cudaEventCreateWithFlags(&dev->ev_render_pos, cudaEventDisableTiming));
            dev->overlap_stages = getenv("WGB_NO_OVERLAP") == nullptr;
            CUDA_CHECK(cudaStreamCreateWithFlags(&dev->copy_stream, cudaStreamNonBlocking));
            dev->no_direct_bins = getenv("WGB_NO_DIRECT_BINS") != nullptr;      // testing knob: always count / scan / fill
            dev->small_work_buffers = getenv("WGB_TEST_SMALL_WORK_BUFFERS") != nullptr;
            dev->async_submit = getenv("WGB_SYNC_SUBMIT") == nullptr;                // WGB_SYNC_SUBMIT=1: every draw is waited for where it is issued
            for (auto& e2 : dev->ev) CUDA_CHECK(cudaEventCreate(&e2));
            for (auto& e2 : dev->timer_ev) CUDA_CHECK(cudaEventCreate(&e2));
            CUDA_CHECK(cudaMallocHost((void**)&dev->host_counters, sizeof(WgbCounters)));
        }
        dev->band_rank = dd.band_rank;
        dev->band_count = dd.band_count ? dd.band_count : 1;
        const uint32_t known = WGB_FEATURE_VIEWPORT_DEPTH_RANGE | WGB_FEATURE_COLOR_WRITE_MASK | WGB_FEATURE_SRGB_ENCODE | WGB_FEATURE_DYNAMIC_OFFSETS |
                               WGB_FEATURE_BLEND;
        if (dd.features & ~known) fail(WGB_ERROR_UNSUPPORTED, "unknown device feature bits 0x%x", dd.features & ~known);
        dev->features = dd.features;
        REQUIRE(dev->band_rank < dev->band_count, "band_rank %u >= band_count %u", dev->band_rank, dev->band_count);
        Queue* q = new Queue();
        q->device = dev;
        *out_queue = to_handle<wgb_queue>(q);
        dev->rc.fetch_add(1);
        *out_device = to_handle<wgb_device>(dev.get());
    });
}

wgb_status wgb_device_poll(wgb_device device, int32_t wait, uint64_t submission_index, uint64_t timeout_ns, int32_t* out_poll) {
    return guarded([&] {
        Device* dev = from_handle<Device>(device, "device");
        std::lock_guard<std::recursive_mutex> lk(dev->mu);
        int32_t result = WGB_POLL_OK;
        if (!dev->compile_only) {
            dev->make_current();
            // retire completed submissions in order
            auto retire = [&] {
                while (!dev->inflight.empty() && cudaEventQuery(dev->inflight.front().done) == cudaSuccess) {
                    cudaEventDestroy(dev->inflight.front().done);
                    dev->inflight.pop_front();
                }
            };
            retire();
            if (wait) {                                                                 // device.rs:258-289
                if (dev->inflight.empty()) result = WGB_POLL_QUEUE_EMPTY;
                else {
                    const auto t0 = std::chrono::steady_clock::now();
                    for (;;) {
                        bool done;
                        if (submission_index == WGB_SUBMISSION_ANY) done = false;
                        else {
                            done = true;
                            for (const auto& f : dev->inflight) if (f.index == submission_index) done = false;
                        }
                        if (done) break;
                        if (dev->inflight.empty()) break;
                        if (timeout_ns == 0 || timeout_ns == UINT64_MAX) {
                            CUDA_CHECK(cudaEventSynchronize(dev->inflight.front().done));
                        } else {
                            const auto el = std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count();
                            if ((uint64_t)el > timeout_ns) { result = WGB_POLL_TIMEOUT; break; }
                        }
                        const size_t before = dev->inflight.size();
                        retire();
                        if (submission_index == WGB_SUBMISSION_ANY && dev->inflight.size() < before) break;
                    }
                }
            }
            // what the asynchronous passes left behind: looked at once nothing of them is in flight any more (a failed
            // pass is re-run here, an error it raised is reported below -- where the reference's engine-thread panic
            // would surface, device.rs:498-503)
            if (dev->inflight.empty()) settle(dev);
        } else if (wait) result = WGB_POLL_QUEUE_EMPTY;
        if (out_poll) *out_poll = result;
        if (dev->deferred_status != WGB_OK) {
            const wgb_status s = dev->deferred_status;
            const std::string m = dev->deferred_error;
            dev->deferred_status = WGB_OK;
            dev->deferred_error.clear();
            throw Error(s, m);
        }
    });
}

// ---- buffers ----
wgb_status wgb_device_create_buffer(wgb_device device, const wgb_buffer_descriptor* desc, wgb_buffer* out) {
    return guarded([&] {
        Device* dev = from_handle<Device>(device, "device");
        REQUIRE(desc && out, "null argument");
        Ref<Buffer> b;
        b.p = new Buffer();
        b->device = Ref<Device>(dev);
        b->size = desc->size;
        b->usage = desc->usage;
        if (!dev->compile_only) {
            dev->make_current();
            CUDA_CHECK(cudaMalloc(&b->dptr, std::max<uint64_t>(desc->size, 16)));
            CUDA_CHECK(cudaMemsetAsync(b->dptr, 0, std::max<uint64_t>(desc->size, 16), dev->joined()));   // Vec<u8> is zero-initialised (buffer.rs:31-35)
        }
        if (desc->mapped_at_creation) {
            b->staging.assign(desc->size, 0);
            b->mapped = true; b->map_mode = WGB_MAP_MODE_WRITE; b->map_offset = 0; b->map_size = desc->size;
        }
        b->rc.fetch_add(1);
        *out = to_handle<wgb_buffer>(b.get());
    });
}
wgb_status wgb_buffer_map_async(wgb_buffer buffer, uint32_t mode, uint64_t offset, uint64_t size, wgb_map_callback callback, void* userdata) {
    wgb_status s = guarded([&] {
        Buffer* b = from_handle<Buffer>(buffer, "buffer");
        std::lock_guard<std::mutex> lk(b->mu);
        REQUIRE(!b->mapped, "buffer is already mapped");
        REQUIRE(mode == WGB_MAP_MODE_READ || mode == WGB_MAP_MODE_WRITE, "invalid map mode");
        if (size == WGB_WHOLE_SIZE) size = b->size - std::min(offset, b->size);
        REQUIRE(range_ok(offset, size, b->size), "map range out of bounds");
        b->staging.assign(b->size, 0);
        Device* dev = b->device.get();
        if (!dev->compile_only) {
            std::lock_guard<std::recursive_mutex> dl(dev->mu);
            dev->make_current();
        settle(dev);
            b->acquire_on(dev->joined());
            CUDA_CHECK(cudaStreamSynchronize(dev->joined()));
            // both modes start from the buffer's contents (a write guard derefs to the live Vec<u8>)
            if (b->size) CUDA_CHECK(cudaMemcpy(b->staging.data(), b->dptr, b->size, cudaMemcpyDeviceToHost));
        }
        b->mapped = true; b->map_mode = mode; b->map_offset = offset; b->map_size = size;
    });
    if (callback) callback(s, userdata);
    return s;
}
wgb_status wgb_buffer_get_mapped_range(wgb_buffer buffer, uint64_t offset, uint64_t size, void** out_ptr) {
    return guarded([&] {
        Buffer* b = from_handle<Buffer>(buffer, "buffer");
        REQUIRE(out_ptr, "out is null");
        std::lock_guard<std::mutex> lk(b->mu);
        REQUIRE(b->mapped, "buffer is not mapped");
        if (size == WGB_WHOLE_SIZE) size = (b->map_offset + b->map_size) - std::min(offset, b->map_offset + b->map_size);
        REQUIRE(offset >= b->map_offset && range_ok(offset, size, b->map_offset + b->map_size), "range lies outside the mapped range");
        *out_ptr = b->staging.data() + offset;
    });
}
wgb_status wgb_buffer_unmap(wgb_buffer buffer) {
    return guarded([&] {
        Buffer* b = from_handle<Buffer>(buffer, "buffer");
        std::lock_guard<std::mutex> lk(b->mu);
        if (!b->mapped) return;
        Device* dev = b->device.get();
        if (b->map_mode == WGB_MAP_MODE_WRITE && b->size && !dev->compile_only) {
            std::lock_guard<std::recursive_mutex> dl(dev->mu);
            dev->make_current();
            settle(dev);
            dev->forget_base_writes(b);
            b->acquire_on(dev->joined());
            CUDA_CHECK(cudaMemcpyAsync(b->dptr, b->staging.data(), b->size, cudaMemcpyHostToDevice, dev->joined()));
            CUDA_CHECK(cudaStreamSynchronize(dev->joined()));
            b->mark_used(dev->joined());
        }
        b->mapped = false;
        std::vector<uint8_t>().swap(b->staging);
    });
}
static void write_buffer_impl(wgb_queue queue, wgb_buffer buffer, uint64_t offset, const void* data, uint64_t size, bool async) {
    Queue* q = from_handle<Queue>(queue, "queue");
    Buffer* b = from_handle<Buffer>(buffer, "buffer");
    REQUIRE(data || size == 0, "data is null");
    REQUIRE(range_ok(offset, size, b->size), "write_buffer range out of bounds");
    Device* dev = q->device.get();
    REQUIRE(b->device.get() == dev, "buffer belongs to another device");
    if (dev->compile_only || size == 0) return;
    std::lock_guard<std::recursive_mutex> dl(dev->mu);
    dev->make_current();
    const bool small = size <= (64u << 10) && !b->external_writers;
    if (small && !dev->unsettled.empty()) {
        // (a region is queued behind passes only with its contents at the head of the log known, and no other extent of
        // it in the log -- Device::base_writes)
        bool known = dev->find_base_write(b, offset, size) != nullptr;
        for (const auto& u : dev->unsettled)
            if (u.write_buffer == b && !(u.write_offset == offset && u.write_data.size() == size) &&
                u.write_offset < offset + size && offset < u.write_offset + u.write_data.size()) known = false;
        if (!known) settle(dev);
    }
    if (small && dev->unsettled.empty()) dev->note_base_write(b, offset, data, size);
    if (!small) dev->forget_base_writes(b);
    if (!dev->unsettled.empty()) {
        // passes are in flight that may have to run again (Device::unsettled): a small write is remembered so that the
        // re-run sees the buffer as it was at that point of the queue; a large one waits for them if they use the buffer
        if (small) {
            dev->unsettled.emplace_back();
            Device::Unsettled& u = dev->unsettled.back();
            u.write_buffer = b;
            u.write_ref = std::make_shared<Ref<Buffer>>(b);
            u.write_offset = offset;
            u.write_data.assign((const uint8_t*)data, (const uint8_t*)data + size);
        } else if (b->last_use_submission != 0) settle(dev);
    }
    cudaPointerAttributes attr;
    const bool pinned = cudaPointerGetAttributes(&attr, data) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    cudaGetLastError();
    if (async) REQUIRE(pinned, "wgb_queue_write_buffer_pinned_async needs page-locked host memory");
    if (pinned && (async || size >= (1u << 20))) {
        // an upload from pinned memory: DMA on the copy stream, ordered after the last use of this buffer only, so it
        // overlaps rendering that reads other buffers; the next submission that uses the buffer waits for it
        if (b->use_pending) { CUDA_CHECK(cudaStreamWaitEvent(dev->copy_stream, b->ev_use, 0)); b->use_pending = false; }
        if (b->write_pending) CUDA_CHECK(cudaStreamWaitEvent(dev->copy_stream, b->ev_write, 0));
        CUDA_CHECK(cudaMemcpyAsync((char*)b->dptr + offset, data, size, cudaMemcpyHostToDevice, dev->copy_stream));
        if (!b->ev_write) CUDA_CHECK(cudaEventCreateWithFlags(&b->ev_write, cudaEventDisableTiming));
        CUDA_CHECK(cudaEventRecord(b->ev_write, dev->copy_stream));
        b->write_pending = true;
        // wgpu's contract (device.rs:332-344): `data` may be reused as soon as write_buffer returns
        if (!async) CUDA_CHECK(cudaEventSynchronize(b->ev_write));
    } else {
        // pageable source: the runtime stages the bytes before cudaMemcpyAsync returns, so `data` may be reused
        // (behind the tile kernels only where one of them may read this buffer in place -- Device::draw_serial)
        cudaStream_t ws = b->last_live_draw > dev->joined_draw_serial || size > (64u << 10) ? dev->joined() : dev->stream;
        b->acquire_on(ws);
        CUDA_CHECK(cudaMemcpyAsync((char*)b->dptr + offset, data, size, cudaMemcpyHostToDevice, ws));
        b->mark_used(ws);
    }
}
wgb_status wgb_queue_write_buffer(wgb_queue queue, wgb_buffer buffer, uint64_t offset, const void* data, uint64_t size) {
    return guarded([&] { write_buffer_impl(queue, buffer, offset, data, size, false); });
}
wgb_status wgb_queue_write_buffer_pinned_async(wgb_queue queue, wgb_buffer buffer, uint64_t offset, const void* data, uint64_t size) {
    return guarded([&] { write_buffer_impl(queue, buffer, offset, data, size, true); });
}
wgb_status wgb_queue_wait_uploads(wgb_queue queue) {
    return guarded([&] {
        Queue* q = from_handle<Queue>(queue, "queue");
        Device* dev = q->device.get();
        if (dev->compile_only) return;
        std::lock_guard<std::recursive_mutex> dl(dev->mu);
        dev->make_current();
        CUDA_CHECK(cudaStreamSynchronize(dev->copy_stream));
    });
}
wgb_status wgb_buffer_device_pointer(wgb_buffer buffer, uint64_t* out_ptr, uint64_t* out_size) {
    return guarded([&] {
        Buffer* b = from_handle<Buffer>(buffer, "buffer");
        REQUIRE(out_ptr, "out is null");
        {
            std::lock_guard<std::recursive_mutex> dl(b->device->mu);
            b->external_writers = true;
            b->device->forget_base_writes(b);
        }
        *out_ptr = (uint64_t)(uintptr_t)b->dptr;
        if (out_size) *out_size = b->size;
    });
}

// ---- textures / samplers ----
wgb_status wgb_device_create_texture(wgb_device device, const wgb_texture_descriptor* desc, wgb_texture* out) {
    return guarded([&] {
        Device* dev = from_handle<Device>(device, "device");
        REQUIRE(desc && out, "null argument");
        const uint32_t bpp = bytes_per_texel(desc->format);
        if (bpp == 0) fail(WGB_ERROR_UNSUPPORTED, "Unsupported texture format: %u", desc->format);   // texture.rs:262-263
        REQUIRE(desc->width > 0 && desc->height > 0, "texture extent is zero");
        REQUIRE(desc->width <= 16384 && desc->height <= 16384, "texture extent exceeds 16384");
        Ref<Texture> t;
        t.p = new Texture();
        t->device = Ref<Device>(dev);
        t->desc = *desc;
        if (t->desc.depth_or_array_layers == 0) t->desc.depth_or_array_layers = 1;
        t->bpp = bpp;
        t->size = (uint64_t)bpp * desc->width * desc->height * t->desc.depth_or_array_layers;
        if (!dev->compile_only) {
            dev->make_current();
            CUDA_CHECK(cudaMalloc(&t->dptr, t->size));
            CUDA_CHECK(cudaMemsetAsync(t->dptr, 0, t->size, dev->joined()));
        }
        t->rc.fetch_add(1);
        *out = to_handle<wgb_texture>(t.get());
    });
}
wgb_status wgb_texture_create_view(wgb_texture texture, const wgb_texture_view_descriptor* desc, wgb_texture_view* out) {
    return guarded([&] {
        Texture* t = from_handle<Texture>(texture, "texture");
        REQUIRE(out, "out is null");
        TextureView* v = new TextureView();
        v->texture = Ref<Texture>(t);
        v->base_layer = desc ? desc->base_array_layer : 0;
        if (v->base_layer >= t->desc.depth_or_array_layers) { delete v; fail(WGB_ERROR_VALIDATION, "base_array_layer out of range"); }
        *out = to_handle<wgb_texture_view>(v);
    });
}
wgb_status wgb_queue_write_texture(wgb_queue queue, wgb_texture texture, uint32_t x, uint32_t y, const void* data, uint64_t data_size,
                                   uint32_t bytes_per_row, uint32_t width, uint32_t height) {
    return guarded([&] {
        Queue* q = from_handle<Queue>(queue, "queue");
        Texture* t = from_handle<Texture>(texture, "texture");
        REQUIRE(data, "data is null");
        REQUIRE((uint64_t)x + width <= t->desc.width && (uint64_t)y + height <= t->desc.height, "write_texture rectangle out of bounds");
        const uint64_t row = (uint64_t)width * t->bpp;
        const uint64_t pitch = bytes_per_row ? bytes_per_row : row;
        REQUIRE(pitch >= row, "bytes_per_row smaller than a row");
        REQUIRE(height == 0 || (uint64_t)(height - 1) * pitch + row <= data_size, "write_texture source too small");
        Device* dev = q->device.get();
        if (dev->compile_only || width == 0 || height == 0) return;
        std::lock_guard<std::recursive_mutex> dl(dev->mu);
        dev->make_current();
        settle(dev);
        t->acquire_for_write(dev->joined());
        char* dst = (char*)t->dptr + ((uint64_t)y * t->desc.width + x) * t->bpp;
        CUDA_CHECK(cudaMemcpy2DAsync(dst, (size_t)t->desc.width * t->bpp, data, pitch, row, height, cudaMemcpyHostToDevice, dev->joined()));
        CUDA_CHECK(cudaStreamSynchronize(dev->joined()));
    });
}
wgb_status wgb_texture_read(wgb_texture texture, void* dst, uint64_t dst_size) {
    return guarded([&] {
        Texture* t = from_handle<Texture>(texture, "texture");
        REQUIRE(dst && dst_size >= t->size, "destination too small: need %llu bytes", (unsigned long long)t->size);
        Device* dev = t->device.get();
        if (dev->compile_only) fail(WGB_ERROR_DEVICE, "a compile-only device has no texture storage");
        std::lock_guard<std::recursive_mutex> dl(dev->mu);
        dev->make_current();
        settle(dev);
        CUDA_CHECK(cudaMemcpyAsync(dst, t->dptr, t->size, cudaMemcpyDeviceToHost, dev->joined()));
        CUDA_CHECK(cudaStreamSynchronize(dev->joined()));
    });
}
// The read-back that is not waited for: the texels as every submission made so far leaves them, copied to page-locked
// host memory on the device's read-back stream, so that the copy overlaps the submissions that follow (they wait for it
// only where they write this texture).  `dst` is valid after wgb_device_wait_readbacks.
wgb_status wgb_texture_read_pinned_async(wgb_texture texture, void* dst, uint64_t dst_size) {
    return guarded([&] {
        Texture* t = from_handle<Texture>(texture, "texture");
        REQUIRE(dst && dst_size >= t->size, "destination too small: need %llu bytes", (unsigned long long)t->size);
        Device* dev = t->device.get();
        if (dev->compile_only) fail(WGB_ERROR_DEVICE, "a compile-only device has no texture storage");
        std::lock_guard<std::recursive_mutex> dl(dev->mu);
        dev->make_current();
        cudaPointerAttributes attr;
        const bool pinned = cudaPointerGetAttributes(&attr, dst) == cudaSuccess && attr.type == cudaMemoryTypeHost;
        cudaGetLastError();
        REQUIRE(pinned, "wgb_texture_read_pinned_async needs page-locked host memory");
        settle(dev);            // passes that have to run again have done so before their target is read
        if (!dev->readback_stream) {
            CUDA_CHECK(cudaStreamCreateWithFlags(&dev->readback_stream, cudaStreamNonBlocking));
            CUDA_CHECK(cudaEventCreateWithFlags(&dev->ev_readback_order, cudaEventDisableTiming));
        }
        CUDA_CHECK(cudaEventRecord(dev->ev_readback_order, dev->joined()));
        CUDA_CHECK(cudaStreamWaitEvent(dev->readback_stream, dev->ev_readback_order, 0));
        CUDA_CHECK(cudaMemcpyAsync(dst, t->dptr, t->size, cudaMemcpyDeviceToHost, dev->readback_stream));
        if (!t->ev_read) CUDA_CHECK(cudaEventCreateWithFlags(&t->ev_read, cudaEventDisableTiming));
        CUDA_CHECK(cudaEventRecord(t->ev_read, dev->readback_stream));
        t->read_pending = true;
    });
}
wgb_status wgb_device_wait_readbacks(wgb_device device) {
    return guarded([&] {
        Device* dev = from_handle<Device>(device, "device");
        if (dev->compile_only) return;
        std::lock_guard<std::recursive_mutex> dl(dev->mu);
        dev->make_current();
        if (dev->readback_stream) CUDA_CHECK(cudaStreamSynchronize(dev->readback_stream));
    });
}
wgb_status wgb_write_png(const char* path, const void* pixels, uint32_t width, uint32_t height, uint32_t channels) {
    return guarded([&] { write_png(path, static_cast<const uint8_t*>(pixels), width, height, channels); });
}
wgb_status wgb_texture_dump_png(wgb_texture texture, const char* path) {
    return guarded([&] {
        Texture* t = from_handle<Texture>(texture, "texture");
        REQUIRE(path, "path is null");
        Device* dev = t->device.get();
        if (dev->compile_only) fail(WGB_ERROR_DEVICE, "a compile-only device has no texture storage");
        const uint32_t w = t->desc.width, h = t->desc.height;
        std::vector<uint8_t> texels((size_t)w * h * t->bpp);
        {
            std::lock_guard<std::recursive_mutex> dl(dev->mu);
            dev->make_current();
            settle(dev);
            CUDA_CHECK(cudaMemcpyAsync(texels.data(), t->dptr, texels.size(), cudaMemcpyDeviceToHost, dev->joined()));
            CUDA_CHECK(cudaStreamSynchronize(dev->joined()));
        }
        switch (t->desc.format) {                                                       // lib.rs:124-154
            case WGB_TEXTURE_FORMAT_RGBA8_UNORM: case WGB_TEXTURE_FORMAT_RGBA8_UNORM_SRGB:
                write_png(path, texels.data(), w, h, 4);
                break;
            case WGB_TEXTURE_FORMAT_BGRA8_UNORM: case WGB_TEXTURE_FORMAT_BGRA8_UNORM_SRGB:
                for (size_t i = 0; i < texels.size(); i += 4) std::swap(texels[i], texels[i + 2]);
                write_png(path, texels.data(), w, h, 4);
                break;
            case WGB_TEXTURE_FORMAT_DEPTH32_FLOAT: {
                std::vector<uint8_t> grey((size_t)w * h);
                for (size_t i = 0; i < grey.size(); i++) {
                    float d;
                    memcpy(&d, texels.data() + i * 4, 4);
                    const float v = d * 255.0f;                                         // `as u8`: saturating, NaN -> 0
                    grey[i] = !(v == v) ? 0 : v <= 0.0f ? 0 : v >= 255.0f ? 255 : (uint8_t)v;
                }
                write_png(path, grey.data(), w, h, 1);
                break;
            }
            default: fail(WGB_ERROR_UNSUPPORTED, "output texture format not implemented: %u (lib.rs:155)", t->desc.format);
        }
    });
}
wgb_status wgb_texture_export_ipc(wgb_texture texture, uint8_t handle[WGB_IPC_HANDLE_SIZE]) {
    return guarded([&] {
        Texture* t = from_handle<Texture>(texture, "texture");
        REQUIRE(handle, "handle is null");
        static_assert(sizeof(cudaIpcMemHandle_t) == WGB_IPC_HANDLE_SIZE, "IPC handle size");
        if (t->device->compile_only) fail(WGB_ERROR_DEVICE, "a compile-only device has no texture storage");
        REQUIRE(!t->imported, "an imported texture cannot be exported again");
        t->device->make_current();
        cudaIpcMemHandle_t h;
        CUDA_CHECK(cudaIpcGetMemHandle(&h, t->dptr));
        memcpy(handle, &h, sizeof(h));
    });
}
wgb_status wgb_device_import_texture_ipc(wgb_device device, const uint8_t handle[WGB_IPC_HANDLE_SIZE],
                                         const wgb_texture_descriptor* desc, wgb_texture* out) {
    return guarded([&] {
        Device* dev = from_handle<Device>(device, "device");
        REQUIRE(handle && desc && out, "null argument");
        if (dev->compile_only) fail(WGB_ERROR_DEVICE, "a compile-only device cannot map peer memory");
        const uint32_t bpp = bytes_per_texel(desc->format);
        if (bpp == 0) fail(WGB_ERROR_UNSUPPORTED, "Unsupported texture format: %u", desc->format);
        Ref<Texture> t;
        t.p = new Texture();
        t->device = Ref<Device>(dev);
        t->desc = *desc;
        if (t->desc.depth_or_array_layers == 0) t->desc.depth_or_array_layers = 1;
        t->bpp = bpp;
        t->size = (uint64_t)bpp * desc->width * desc->height * t->desc.depth_or_array_layers;
        dev->make_current();
        cudaIpcMemHandle_t h;
        memcpy(&h, handle, sizeof(h));
        CUDA_CHECK(cudaIpcOpenMemHandle(&t->dptr, h, cudaIpcMemLazyEnablePeerAccess));
        t->imported = true;
        t->rc.fetch_add(1);
        *out = to_handle<wgb_texture>(t.get());
    });
}
wgb_status wgb_texture_device_pointer(wgb_texture texture, uint64_t* out_ptr, uint64_t* out_size) {
    return guarded([&] {
        Texture* t = from_handle<Texture>(texture, "texture");
        if (out_ptr) *out_ptr = (uint64_t)(uintptr_t)t->dptr;
        if (out_size) *out_size = t->size;
    });
}
wgb_status wgb_device_create_sampler(wgb_device device, const wgb_sampler_descriptor* desc, wgb_sampler* out) {
    return guarded([&] {
        from_handle<Device>(device, "device");
        REQUIRE(desc && out, "null argument");
        if (desc->address_mode_u == WGB_ADDRESS_MODE_CLAMP_TO_BORDER || desc->address_mode_v == WGB_ADDRESS_MODE_CLAMP_TO_BORDER)
            fail(WGB_ERROR_UNSUPPORTED, "ClampToBorder (binding.rs:161 todo!)");
        REQUIRE(desc->address_mode_u <= 2 && desc->address_mode_v <= 2, "invalid address mode");
        Sampler* s = new Sampler();
        s->desc = *desc;      // filters are kept but sampling is always nearest, like the reference (binding.rs:93-149)
        *out = to_handle<wgb_sampler>(s);
    });
}

// ---- shaders ----
wgb_status wgb_device_create_shader_module(wgb_device device, const wgb_shader_module_descriptor* desc, wgb_shader_module* out) {
    return guarded([&] {
        from_handle<Device>(device, "device");
        REQUIRE(desc && out, "null argument");
        REQUIRE(desc->wgsl || desc->emitted_count, "shader module needs WGSL source or emitted entry points");
        std::unique_ptr<ShaderModule> m(new ShaderModule());
        if (desc->wgsl) m->wgsl = desc->wgsl;
        for (uint32_t i = 0; i < desc->emitted_count; i++) {
            REQUIRE(desc->emitted[i].entry_point && desc->emitted[i].cuda_source, "emitted entry point is incomplete");
            m->emitted.push_back({desc->emitted[i].stage, desc->emitted[i].entry_point, desc->emitted[i].cuda_source});
        }
        *out = to_handle<wgb_shader_module>(m.release());
    });
}
wgb_status wgb_translate_wgsl(const char* wgsl, uint32_t stage, const char* entry_point, char** out_cuda) {
    return guarded([&] {
        REQUIRE(wgsl && entry_point && out_cuda, "null argument");
        const std::string s = emit_wgsl(wgsl, stage, entry_point);
        *out_cuda = (char*)malloc(s.size() + 1);
        if (!*out_cuda) throw std::bad_alloc();
        memcpy(*out_cuda, s.c_str(), s.size() + 1);
    });
}

// ---- binding model ----
wgb_status wgb_device_create_bind_group_layout(wgb_device device, const wgb_bind_group_layout_entry* entries, uint32_t count, wgb_bind_group_layout* out) {
    return guarded([&] {
        from_handle<Device>(device, "device");
        REQUIRE(out && (entries || count == 0), "null argument");
        BindGroupLayout* l = new BindGroupLayout();
        l->entries.assign(entries, entries + count);
        *out = to_handle<wgb_bind_group_layout>(l);
    });
}
wgb_status wgb_device_create_pipeline_layout(wgb_device device, const wgb_bind_group_layout* layouts, uint32_t count, wgb_pipeline_layout* out) {
    return guarded([&] {
        from_handle<Device>(device, "device");
        REQUIRE(out && (layouts || count == 0), "null argument");
        std::unique_ptr<PipelineLayout> l(new PipelineLayout());
        for (uint32_t i = 0; i < count; i++) l->layouts.push_back(Ref<BindGroupLayout>(from_handle<BindGroupLayout>(layouts[i], "bind group layout")));
        *out = to_handle<wgb_pipeline_layout>(l.release());
    });
}
wgb_status wgb_device_create_bind_group(wgb_device device, wgb_bind_group_layout layout, const wgb_bind_group_entry* entries, uint32_t count, wgb_bind_group* out) {
    return guarded([&] {
        from_handle<Device>(device, "device");
        if (layout) from_handle<BindGroupLayout>(layout, "bind group layout");
        REQUIRE(out && (entries || count == 0), "null argument");
        std::unique_ptr<BindGroup> g(new BindGroup());
        if (layout) g->layout = Ref<BindGroupLayout>(from_handle<BindGroupLayout>(layout, "bind group layout"));
        for (uint32_t i = 0; i < count; i++) {
            BindGroup::Entry e;
            e.binding = entries[i].binding;
            e.kind = entries[i].kind;
            if (e.kind == WGB_BINDING_BUFFER) {
                Buffer* b = from_handle<Buffer>(entries[i].buffer, "buffer");
                e.buffer = Ref<Buffer>(b);
                e.offset = entries[i].offset;
                e.size = entries[i].size == WGB_WHOLE_SIZE ? b->size - std::min(b->size, e.offset) : entries[i].size;
            } else if (e.kind == WGB_BINDING_TEXTURE_VIEW) e.view = Ref<TextureView>(from_handle<TextureView>(entries[i].texture_view, "texture view"));
            else if (e.kind == WGB_BINDING_SAMPLER) e.sampler = Ref<Sampler>(from_handle<Sampler>(entries[i].sampler, "sampler"));
            else fail(WGB_ERROR_VALIDATION, "invalid binding kind %u", e.kind);
            g->entries.push_back(e);
        }
        *out = to_handle<wgb_bind_group>(g.release());
    });
}

// ---- render pipeline ----
wgb_status wgb_device_create_render_pipeline(wgb_device device, const wgb_render_pipeline_descriptor* desc, wgb_render_pipeline* out) {
    return guarded([&] {
        Device* dev = from_handle<Device>(device, "device");
        REQUIRE(desc && out, "null argument");
        if (desc->polygon_mode != WGB_POLYGON_MODE_FILL) fail(WGB_ERROR_UNSUPPORTED, "PolygonMode::Line/Point (state.rs:438-478 panics)");
        REQUIRE(desc->topology <= WGB_TOPOLOGY_TRIANGLE_STRIP, "invalid topology");
        Ref<RenderPipeline> p;
        p.p = new RenderPipeline();
        p->device = Ref<Device>(dev);
        ShaderModule* vm = from_handle<ShaderModule>(desc->vertex_module, "vertex shader module");
        p->vs_text = vm->cuda_for(WGB_SHADER_STAGE_VERTEX, desc->vertex_entry_point ? desc->vertex_entry_point : "vs_main");
        if (desc->fragment_module) {
            ShaderModule* fm = from_handle<ShaderModule>(desc->fragment_module, "fragment shader module");
            p->fs_text = fm->cuda_for(WGB_SHADER_STAGE_FRAGMENT, desc->fragment_entry_point ? desc->fragment_entry_point : "fs_main");
            p->has_fragment = true;
        }
        REQUIRE(desc->vertex_buffer_count <= WGB_MAX_VERTEX_BUFFERS, "too many vertex buffers");
        for (uint32_t b = 0; b < desc->vertex_buffer_count; b++) {
            RenderPipeline::VB vb;
            vb.stride = desc->vertex_buffers[b].array_stride;
            vb.step_mode = desc->vertex_buffers[b].step_mode;
            for (uint32_t a = 0; a < desc->vertex_buffers[b].attribute_count; a++) {
                const wgb_vertex_attribute& at = desc->vertex_buffers[b].attributes[a];
                REQUIRE(vertex_format_size(at.format) != 0, "unsupported vertex format %u", at.format);
                REQUIRE(at.shader_location < 32, "shader_location too large");
                vb.attrs.push_back(at);
            }
            p->vbs.push_back(vb);
        }
        p->topology = desc->topology; p->strip_index_format = desc->strip_index_format;
        p->front_face = desc->front_face; p->cull_mode = desc->cull_mode;
        p->has_depth_stencil = desc->has_depth_stencil != 0;
        p->depth_format = desc->depth_format; p->depth_write = desc->depth_write_enabled; p->depth_compare = desc->depth_compare;
        if (p->has_depth_stencil) REQUIRE(desc->depth_compare >= 1 && desc->depth_compare <= 8, "invalid depth compare");
        REQUIRE(desc->target_count <= WGB_MAX_COLOR, "too many colour targets");
        for (uint32_t t = 0; t < desc->target_count; t++) {
            REQUIRE(is_color_format(desc->targets[t].format), "colour target %u has a non-colour format", t);
            p->targets.push_back(desc->targets[t]);   // blend and write_mask are accepted and not applied (fragment.rs:480-485)
        }
        // compile the variant the descriptor implies now (device.rs:129-134 compiles at creation)
        p->variant(p->has_depth_stencil, false);
        p->rc.fetch_add(1);
        *out = to_handle<wgb_render_pipeline>(p.get());
    });
}
wgb_status wgb_render_pipeline_get_source(wgb_render_pipeline pipeline, char** out) {
    return guarded([&] {
        RenderPipeline* p = from_handle<RenderPipeline>(pipeline, "render pipeline");
        REQUIRE(out, "out is null");
        std::lock_guard<std::mutex> lk(p->mu);
        REQUIRE(!p->variant_source.empty(), "pipeline has no compiled variant");
        const std::string& s = p->variant_source.begin()->second;
        *out = (char*)malloc(s.size() + 1);
        if (!*out) throw std::bad_alloc();
        memcpy(*out, s.c_str(), s.size() + 1);
    });
}

// ---- command encoding ----
wgb_status wgb_device_create_command_encoder(wgb_device device, wgb_command_encoder* out) {
    return guarded([&] {
        Device* dev = from_handle<Device>(device, "device");
        REQUIRE(out, "out is null");
        CommandEncoder* e = new CommandEncoder();
        e->device = Ref<Device>(dev);
        *out = to_handle<wgb_command_encoder>(e);
    });
}
wgb_status wgb_command_encoder_begin_render_pass(wgb_command_encoder encoder, const wgb_render_pass_descriptor* desc, wgb_render_pass* out) {
    return guarded([&] {
        CommandEncoder* enc = from_handle<CommandEncoder>(encoder, "command encoder");
        REQUIRE(desc && out, "null argument");
        REQUIRE(!enc->finished, "command encoder is finished");
        auto cmd = std::make_shared<PassCommand>();
        for (uint32_t i = 0; i < desc->color_attachment_count; i++) {
            const wgb_color_attachment& a = desc->color_attachments[i];
            PassCommand::Color c;
            if (a.view) c.view = Ref<TextureView>(from_handle<TextureView>(a.view, "texture view"));
            c.load_op = a.load_op; c.store_op = a.store_op;
            memcpy(c.clear, a.clear_value, sizeof(c.clear));
            cmd->colors.push_back(c);
        }
        if (desc->depth_stencil_attachment) {
            const wgb_depth_stencil_attachment& a = *desc->depth_stencil_attachment;
            cmd->has_depth = true;
            cmd->depth_view = Ref<TextureView>(from_handle<TextureView>(a.view, "texture view"));
            cmd->has_depth_ops = a.has_depth_ops != 0;
            cmd->depth_load_op = a.depth_load_op;
            cmd->depth_clear = a.depth_clear_value;
            cmd->has_stencil_ops = a.has_stencil_ops != 0;
        }
        RenderPass* p = new RenderPass();
        p->encoder = Ref<CommandEncoder>(enc);
        p->cmd = cmd;
        *out = to_handle<wgb_render_pass>(p);
    });
}
static SubCommand& record(wgb_render_pass pass, SubCommand::Kind k) {
    RenderPass* p = from_handle<RenderPass>(pass, "render pass");
    REQUIRE(!p->ended, "render pass has ended");
    p->cmd->sub.emplace_back();
    p->cmd->sub.back().kind = k;
    return p->cmd->sub.back();
}
wgb_status wgb_render_pass_set_pipeline(wgb_render_pass pass, wgb_render_pipeline pipeline) {
    return guarded([&] {
        RenderPipeline* rp = from_handle<RenderPipeline>(pipeline, "render pipeline");
        record(pass, SubCommand::SetPipeline).pipeline = Ref<RenderPipeline>(rp);
    });
}
wgb_status wgb_render_pass_set_bind_group(wgb_render_pass pass, uint32_t index, wgb_bind_group group, const uint32_t* dynamic_offsets, uint32_t dynamic_offset_count) {
    return guarded([&] {
        BindGroup* g = group ? from_handle<BindGroup>(group, "bind group") : nullptr;
        REQUIRE(dynamic_offsets || dynamic_offset_count == 0, "dynamic offsets are null");
        SubCommand& s = record(pass, SubCommand::SetBindGroup);
        s.index = index;
        s.dynamic_offsets.assign(dynamic_offsets, dynamic_offsets + dynamic_offset_count);
        if (g) s.bind_group = Ref<BindGroup>(g);
    });
}
wgb_status wgb_render_pass_set_index_buffer(wgb_render_pass pass, wgb_buffer buffer, uint32_t index_format, uint64_t offset, uint64_t size) {
    return guarded([&] {
        Buffer* b = from_handle<Buffer>(buffer, "buffer");
        REQUIRE(index_format == WGB_INDEX_FORMAT_UINT16 || index_format == WGB_INDEX_FORMAT_UINT32, "invalid index format");
        SubCommand& s = record(pass, SubCommand::SetIndexBuffer);
        s.buffer = Ref<Buffer>(b); s.index_format = index_format; s.offset = offset; s.size = size;
    });
}
wgb_status wgb_render_pass_set_vertex_buffer(wgb_render_pass pass, uint32_t slot, wgb_buffer buffer, uint64_t offset, uint64_t size) {
    return guarded([&] {
        Buffer* b = from_handle<Buffer>(buffer, "buffer");
        SubCommand& s = record(pass, SubCommand::SetVertexBuffer);
        s.buffer = Ref<Buffer>(b); s.index = slot; s.offset = offset; s.size = size;
    });
}
wgb_status wgb_render_pass_set_viewport(wgb_render_pass pass, float x, float y, float width, float height, float min_depth, float max_depth) {
    return guarded([&] {
        SubCommand& s = record(pass, SubCommand::SetViewport);
        s.vp[0] = x; s.vp[1] = y; s.vp[2] = width; s.vp[3] = height; s.vp[4] = min_depth; s.vp[5] = max_depth;
    });
}
wgb_status wgb_render_pass_set_scissor_rect(wgb_render_pass pass, uint32_t x, uint32_t y, uint32_t width, uint32_t height) {
    return guarded([&] {
        SubCommand& s = record(pass, SubCommand::SetScissor);
        s.sc[0] = x; s.sc[1] = y; s.sc[2] = width; s.sc[3] = height;
    });
}
wgb_status wgb_render_pass_set_blend_constant(wgb_render_pass pass, const double* color) {
    return guarded([&] {
        REQUIRE(color, "colour is null");
        SubCommand& s = record(pass, SubCommand::SetBlendConstant);
        for (int k = 0; k < 4; k++) s.blend_constant[k] = (float)color[k];
    });
}
wgb_status wgb_render_pass_set_stencil_reference(wgb_render_pass pass, uint32_t) {
    return guarded([&] { record(pass, SubCommand::SetStencilReference); });
}
wgb_status wgb_render_pass_draw(wgb_render_pass pass, uint32_t first_vertex, uint32_t vertex_count, uint32_t first_instance, uint32_t instance_count) {
    return guarded([&] {
        SubCommand& s = record(pass, SubCommand::Draw);
        s.first = first_vertex; s.count = vertex_count; s.first_instance = first_instance; s.instance_count = instance_count;
    });
}
wgb_status wgb_render_pass_draw_indexed(wgb_render_pass pass, uint32_t first_index, uint32_t index_count, int32_t base_vertex,
                                        uint32_t first_instance, uint32_t instance_count) {
    return guarded([&] {
        SubCommand& s = record(pass, SubCommand::DrawIndexed);
        s.first = first_index; s.count = index_count; s.base_vertex = base_vertex; s.first_instance = first_instance; s.instance_count = instance_count;
    });
}
wgb_status wgb_render_pass_end(wgb_render_pass pass) {
    return guarded([&] { from_handle<RenderPass>(pass, "render pass")->end(); });
}
wgb_status wgb_command_encoder_copy_buffer_to_buffer(wgb_command_encoder encoder, wgb_buffer source, uint64_t source_offset,
                                                     wgb_buffer destination, uint64_t destination_offset, uint64_t size) {
    return guarded([&] {
        auto c = std::make_shared<CopyCommand>();
        c->kind = CopyCommand::BufferToBuffer;
        c->src_buffer = Ref<Buffer>(from_handle<Buffer>(source, "buffer"));
        c->dst_buffer = Ref<Buffer>(from_handle<Buffer>(destination, "buffer"));
        if (size == WGB_WHOLE_SIZE) size = c->src_buffer->size - std::min(c->src_buffer->size, source_offset);
        REQUIRE(range_ok(source_offset, size, c->src_buffer->size) && range_ok(destination_offset, size, c->dst_buffer->size), "copy_buffer_to_buffer range out of bounds");
        c->src_offset = source_offset; c->dst_offset = destination_offset; c->size = size;
        record_copy(encoder, c);
    });
}
wgb_status wgb_command_encoder_copy_buffer_to_texture(wgb_command_encoder encoder, const wgb_texel_copy_buffer_info* source,
                                                      const wgb_texel_copy_texture_info* destination, uint32_t width, uint32_t height) {
    return guarded([&] {
        REQUIRE(source && destination, "null argument");
        auto c = std::make_shared<CopyCommand>();
        c->kind = CopyCommand::BufferToTexture;
        c->src_buffer = Ref<Buffer>(from_handle<Buffer>(source->buffer, "buffer"));
        c->dst_texture = Ref<Texture>(from_handle<Texture>(destination->texture, "texture"));
        check_texture_rect(c->dst_texture.get(), destination->x, destination->y, destination->layer, width, height, "copy_buffer_to_texture");
        check_buffer_rect(c->src_buffer.get(), source->offset, source->bytes_per_row, c->dst_texture->bpp, width, height, "copy_buffer_to_texture");
        c->src_offset = source->offset; c->bytes_per_row = source->bytes_per_row;
        c->dst_x = destination->x; c->dst_y = destination->y; c->dst_layer = destination->layer; c->width = width; c->height = height;
        record_copy(encoder, c);
    });
}
wgb_status wgb_command_encoder_copy_texture_to_buffer(wgb_command_encoder encoder, const wgb_texel_copy_texture_info* source,
                                                      const wgb_texel_copy_buffer_info* destination, uint32_t width, uint32_t height) {
    return guarded([&] {
        REQUIRE(source && destination, "null argument");
        auto c = std::make_shared<CopyCommand>();
        c->kind = CopyCommand::TextureToBuffer;
        c->src_texture = Ref<Texture>(from_handle<Texture>(source->texture, "texture"));
        c->dst_buffer = Ref<Buffer>(from_handle<Buffer>(destination->buffer, "buffer"));
        check_texture_rect(c->src_texture.get(), source->x, source->y, source->layer, width, height, "copy_texture_to_buffer");
        check_buffer_rect(c->dst_buffer.get(), destination->offset, destination->bytes_per_row, c->src_texture->bpp, width, height, "copy_texture_to_buffer");
        c->dst_offset = destination->offset; c->bytes_per_row = destination->bytes_per_row;
        c->src_x = source->x; c->src_y = source->y; c->src_layer = source->layer; c->width = width; c->height = height;
        record_copy(encoder, c);
    });
}
wgb_status wgb_command_encoder_copy_texture_to_texture(wgb_command_encoder encoder, const wgb_texel_copy_texture_info* source,
                                                       const wgb_texel_copy_texture_info* destination, uint32_t width, uint32_t height) {
    return guarded([&] {
        REQUIRE(source && destination, "null argument");
        auto c = std::make_shared<CopyCommand>();
        c->kind = CopyCommand::TextureToTexture;
        c->src_texture = Ref<Texture>(from_handle<Texture>(source->texture, "texture"));
        c->dst_texture = Ref<Texture>(from_handle<Texture>(destination->texture, "texture"));
        REQUIRE(c->src_texture->bpp == c->dst_texture->bpp, "copy_texture_to_texture: texel sizes differ");
        check_texture_rect(c->src_texture.get(), source->x, source->y, source->layer, width, height, "copy_texture_to_texture source");
        check_texture_rect(c->dst_texture.get(), destination->x, destination->y, destination->layer, width, height, "copy_texture_to_texture destination");
        c->src_x = source->x; c->src_y = source->y; c->src_layer = source->layer;
        c->dst_x = destination->x; c->dst_y = destination->y; c->dst_layer = destination->layer; c->width = width; c->height = height;
        record_copy(encoder, c);
    });
}
wgb_status wgb_command_encoder_clear_buffer(wgb_command_encoder encoder, wgb_buffer buffer, uint64_t offset, uint64_t size) {
    return guarded([&] {
        auto c = std::make_shared<CopyCommand>();
        c->kind = CopyCommand::ClearBuffer;
        c->dst_buffer = Ref<Buffer>(from_handle<Buffer>(buffer, "buffer"));
        if (size == WGB_WHOLE_SIZE) size = c->dst_buffer->size - std::min(c->dst_buffer->size, offset);
        REQUIRE(range_ok(offset, size, c->dst_buffer->size), "clear_buffer range out of bounds");
        c->dst_offset = offset; c->size = size;
        record_copy(encoder, c);
    });
}
wgb_status wgb_command_encoder_clear_texture(wgb_command_encoder encoder, wgb_texture texture) {
    return guarded([&] {
        auto c = std::make_shared<CopyCommand>();
        c->kind = CopyCommand::ClearTexture;
        c->dst_texture = Ref<Texture>(from_handle<Texture>(texture, "texture"));
        record_copy(encoder, c);
    });
}
wgb_status wgb_command_encoder_finish(wgb_command_encoder encoder, wgb_command_buffer* out) {
    return guarded([&] {
        CommandEncoder* enc = from_handle<CommandEncoder>(encoder, "command encoder");
        REQUIRE(out, "out is null");
        std::lock_guard<std::mutex> lk(enc->mu);
        REQUIRE(!enc->finished, "command encoder is already finished");
        enc->finished = true;
        CommandBuffer* cb = new CommandBuffer();
        cb->device = enc->device;
        cb->passes.swap(enc->passes);
        *out = to_handle<wgb_command_buffer>(cb);
    });
}
wgb_status wgb_queue_submit(wgb_queue queue, const wgb_command_buffer* command_buffers, uint32_t count, uint64_t* out_submission_index) {
    return guarded([&] {
        Queue* q = from_handle<Queue>(queue, "queue");
        Device* dev = q->device.get();
        REQUIRE(command_buffers || count == 0, "null argument");
        std::lock_guard<std::recursive_mutex> lk(dev->mu);
        if (dev->compile_only) fail(WGB_ERROR_DEVICE, "a compile-only device cannot execute submissions");
        dev->make_current();
        const uint64_t index = dev->next_submission++;                                  // device.rs:443-444
        if (out_submission_index) *out_submission_index = index;
        dev->current_submission = index;
        wgb_status st = WGB_OK;
        std::string msg;
        for (uint32_t i = 0; i < count && st == WGB_OK; i++) {
            CommandBuffer* cb = from_handle<CommandBuffer>(command_buffers[i], "command buffer");
            REQUIRE(!cb->submitted, "command buffer was already submitted");
            REQUIRE(cb->device.get() == dev, "command buffer belongs to another device");
            cb->submitted = true;
            // buffers this command buffer touches: wait for uploads still running on the copy stream
            std::vector<Buffer*> used;
            for (const auto& cmd : cb->passes) {
                if (cmd.pass) {
                    for (const SubCommand& sc : cmd.pass->sub) {
                        if (sc.buffer) used.push_back(sc.buffer.get());
                        if (sc.bind_group) for (const auto& e : sc.bind_group->entries) if (e.buffer) used.push_back(e.buffer.get());
                    }
                } else {
                    if (cmd.copy->src_buffer) used.push_back(cmd.copy->src_buffer.get());
                    if (cmd.copy->dst_buffer) used.push_back(cmd.copy->dst_buffer.get());
                }
            }
            std::sort(used.begin(), used.end());
            used.erase(std::unique(used.begin(), used.end()), used.end());
            // the recorded commands may hold the last reference to a buffer (the application is free to drop its handle
            // once the pass is recorded) and are cleared below, before this scope ends: keep the buffers alive until the
            // end-of-scope bookkeeping has run
            std::vector<Ref<Buffer>> keep_alive(used.begin(), used.end());
            for (Buffer* b : used) { b->acquire_on(dev->stream); b->last_use_submission = index; }      // (waits only)
            struct MarkUsed {
                std::vector<Buffer*>& v; Device* d;
                ~MarkUsed() { if (v.empty()) return; cudaStream_t s = d->tail(); for (Buffer* b : v) b->mark_used(s); }      // behind the tile kernels too
            } mark_used{used, dev};
            for (const auto& cmd : cb->passes) {
                // errors raised while a submission executes surface at poll, where the reference's
                // engine-thread panic would be observed (device.rs:498-503)
                try {
                    if (cmd.pass) execute_pass(dev, cmd.pass, true);
                    else { settle(dev); execute_copy(dev, *cmd.copy); }
                }
                catch (const Error& e) { st = e.status; msg = e.what(); break; }
            }
            cb->passes.clear();
        }
        cudaEvent_t done;
        CUDA_CHECK(cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
        CUDA_CHECK(cudaEventRecord(done, dev->tail()));
        dev->inflight.push_back({index, done});
        if (st != WGB_OK && dev->deferred_status == WGB_OK) { dev->deferred_status = st; dev->deferred_error = msg; }
    });
}

// ---- measurement ----
wgb_status wgb_device_get_last_pass_stats(wgb_device device, wgb_pass_stats* out) {
    return guarded([&] {
        Device* dev = from_handle<Device>(device, "device");
        REQUIRE(out, "out is null");
        std::lock_guard<std::recursive_mutex> lk(dev->mu);
        if (!dev->compile_only) { dev->make_current(); settle(dev); }
        *out = dev->last_stats;
    });
}
wgb_status wgb_device_set_coverage_capture(wgb_device device, int32_t enabled) {
    return guarded([&] {
        Device* dev = from_handle<Device>(device, "device");
        std::lock_guard<std::recursive_mutex> lk(dev->mu);
        if (!dev->compile_only) { dev->make_current(); settle(dev); }
        dev->coverage_capture = enabled != 0;
        dev->coverage_w = dev->coverage_h = 0;
    });
}
wgb_status wgb_device_read_coverage(wgb_device device, uint32_t* dst, uint64_t pixel_count) {
    return guarded([&] {
        Device* dev = from_handle<Device>(device, "device");
        std::lock_guard<std::recursive_mutex> lk(dev->mu);
        REQUIRE(dst, "dst is null");
        REQUIRE(dev->coverage_capture && dev->coverage.p, "coverage capture is not enabled or no pass has run");
        REQUIRE(pixel_count == (uint64_t)dev->coverage_w * dev->coverage_h, "pixel_count does not match the last pass (%ux%u)", dev->coverage_w, dev->coverage_h);
        dev->make_current();
        settle(dev);
        CUDA_CHECK(cudaMemcpyAsync(dst, dev->coverage.p, pixel_count * 4, cudaMemcpyDeviceToHost, dev->joined()));
        CUDA_CHECK(cudaStreamSynchronize(dev->joined()));
    });
}
wgb_status wgb_device_set_band(wgb_device device, uint32_t band_rank, uint32_t band_count) {
    return guarded([&] {
        Device* dev = from_handle<Device>(device, "device");
        std::lock_guard<std::recursive_mutex> lk(dev->mu);
        if (band_count == 0) band_count = 1;
        REQUIRE(band_rank < band_count, "band_rank %u >= band_count %u", band_rank, band_count);
        if (!dev->compile_only) { dev->make_current(); settle(dev); }
        dev->band_rank = band_rank; dev->band_count = band_count;
    });
}
wgb_status wgb_device_get_band_rows(wgb_device device, uint32_t height, uint32_t* out_row0, uint32_t* out_row1) {
    return guarded([&] {
        Device* dev = from_handle<Device>(device, "device");
        std::lock_guard<std::recursive_mutex> lk(dev->mu);
        uint32_t ty0, ty1;
        band_rows(dev, (height + WGB_TILE_H - 1) / WGB_TILE_H, ty0, ty1);
        if (out_row0) *out_row0 = std::min(ty0 * WGB_TILE_H, height);
        if (out_row1) *out_row1 = std::min(ty1 * WGB_TILE_H, height);
    });
}
wgb_status wgb_device_timer_begin(wgb_device device) {
    return guarded([&] {
        Device* dev = from_handle<Device>(device, "device");
        std::lock_guard<std::recursive_mutex> lk(dev->mu);
        if (dev->compile_only) fail(WGB_ERROR_DEVICE, "a compile-only device has no stream");
        dev->make_current();
        CUDA_CHECK(cudaEventRecord(dev->timer_ev[0], dev->joined()));
    });
}
wgb_status wgb_device_timer_end(wgb_device device, float* out_ms) {
    return guarded([&] {
        Device* dev = from_handle<Device>(device, "device");
        REQUIRE(out_ms, "out is null");
        std::lock_guard<std::recursive_mutex> lk(dev->mu);
        if (dev->compile_only) fail(WGB_ERROR_DEVICE, "a compile-only device has no stream");
        dev->make_current();
        CUDA_CHECK(cudaEventRecord(dev->timer_ev[1], dev->joined()));
        CUDA_CHECK(cudaEventSynchronize(dev->timer_ev[1]));
        CUDA_CHECK(cudaEventElapsedTime(out_ms, dev->timer_ev[0], dev->timer_ev[1]));
    });
}
wgb_status wgb_device_get_stream(wgb_device device, void** out_stream) {
    return guarded([&] {
        Device* dev = from_handle<Device>(device, "device");
        REQUIRE(out_stream, "out is null");
        *out_stream = (void*)dev->joined();
    });
}

// ---- surface / present (surface.rs) ----
wgb_status wgb_instance_create_surface(wgb_instance instance, const wgb_surface_target* target, wgb_surface* out) {
    return guarded([&] {
        from_handle<Instance>(instance, "instance");
        REQUIRE(out, "out is null");
        Surface* s = new Surface();
        if (target) s->target = *target;
        *out = to_handle<wgb_surface>(s);
    });
}
wgb_status wgb_adapter_is_surface_supported(wgb_adapter adapter, wgb_surface surface, int32_t* out_supported) {
    return guarded([&] {
        from_handle<Adapter>(adapter, "adapter");
        REQUIRE(out_supported, "out is null");
        // adapter.rs:44-54: `surface.as_custom::<Surface>().is_some()`
        *out_supported = surface && dynamic_cast<Surface*>(reinterpret_cast<Object*>(surface)) ? 1 : 0;
    });
}
wgb_status wgb_surface_get_capabilities(wgb_surface surface, wgb_adapter adapter, wgb_surface_capabilities* out) {
    return guarded([&] {
        from_handle<Surface>(surface, "surface");
        from_handle<Adapter>(adapter, "adapter");
        REQUIRE(out, "out is null");
        memset(out, 0, sizeof(*out));
        out->format_count = 1; out->formats[0] = WGB_TEXTURE_FORMAT_BGRA8_UNORM;               // surface.rs:61, 242
        out->present_mode_count = 1; out->present_modes[0] = WGB_PRESENT_MODE_IMMEDIATE;        // surface.rs:64
        out->alpha_mode_count = 1; out->alpha_modes[0] = WGB_COMPOSITE_ALPHA_MODE_OPAQUE;       // surface.rs:66
        out->usages = WGB_TEXTURE_USAGE_RENDER_ATTACHMENT;                                      // surface.rs:68
    });
}
wgb_status wgb_surface_configure(wgb_surface surface, wgb_device device, const wgb_surface_configuration* config) {
    return guarded([&] {
        Surface* s = from_handle<Surface>(surface, "surface");
        Device* dev = from_handle<Device>(device, "device");
        REQUIRE(config, "config is null");
        // check_surface_config (surface.rs:234-257)
        REQUIRE(config->format == WGB_TEXTURE_FORMAT_BGRA8_UNORM, "Unsupported surface texture format: %u", config->format);
        REQUIRE(config->view_format_count == 0 || config->view_formats, "view_formats is null");
        for (uint32_t i = 0; i < config->view_format_count; i++)
            REQUIRE(config->view_formats[i] == WGB_TEXTURE_FORMAT_BGRA8_UNORM, "Unsupported surface texture format: %u", config->view_formats[i]);
        REQUIRE(config->width != 0, "Surface width must not be zero");         // surface.rs:105
        REQUIRE(config->height != 0, "Surface height must not be zero");       // surface.rs:106
        wgb_texture_descriptor td{};
        td.width = config->width; td.height = config->height; td.depth_or_array_layers = 1;
        td.mip_level_count = 1; td.sample_count = 1; td.format = config->format; td.usage = config->usage;
        wgb_texture th = nullptr;
        const wgb_status st = wgb_device_create_texture(device, &td, &th);      // Texture::new (surface.rs:94-101): zeroed
        if (st != WGB_OK) throw Error(st, g_last_error);
        Ref<Texture> tex;
        tex.p = from_handle<Texture>(th, "texture");      // takes over the handle's reference
        std::lock_guard<std::recursive_mutex> lk(s->mu);
        s->configured = false;
        s->drop_window();
        s->device = Ref<Device>(dev);
        s->texture = tex;
        s->config = *config;
        s->config.view_formats = nullptr; s->config.view_format_count = 0;
        s->window_size = tex->size;
        if (dev->compile_only) {
            s->window = calloc(1, s->window_size);
            REQUIRE(s->window, "out of host memory");
        } else {
            std::lock_guard<std::recursive_mutex> dl(dev->mu);
            dev->make_current();
            CUDA_CHECK(cudaMallocHost(&s->window, s->window_size));     // inner.surface.resize (surface.rs:103-109)
            s->window_pinned = true;
            memset(s->window, 0, s->window_size);
        }
        s->configured = true;
    });
}
wgb_status wgb_surface_get_current_texture(wgb_surface surface, wgb_texture* out, uint32_t* out_status) {
    return guarded([&] {
        Surface* s = from_handle<Surface>(surface, "surface");
        REQUIRE(out, "out is null");
        std::lock_guard<std::recursive_mutex> lk(s->mu);
        REQUIRE(s->configured, "Surface not configured yet");            // surface.rs:127-130
        s->texture->rc.fetch_add(1);                                     // `configured.buffer.clone()`
        *out = to_handle<wgb_texture>(s->texture.get());
        if (out_status) *out_status = WGB_SURFACE_STATUS_GOOD;
    });
}
wgb_status wgb_surface_present(wgb_surface surface) {
    return guarded([&] {
        Surface* s = from_handle<Surface>(surface, "surface");
        std::lock_guard<std::recursive_mutex> lk(s->mu);
        REQUIRE(s->configured, "Surface not configured yet");            // surface.rs:174-177
        Texture* t = s->texture.get();
        Device* dev = s->device.get();
        if (dev->compile_only) fail(WGB_ERROR_DEVICE, "a compile-only device has no texture storage");
        {
            // `wait.wait()` + `buffer.read()` (surface.rs:179-183): everything submitted so far has written the texture
            std::lock_guard<std::recursive_mutex> dl(dev->mu);
            dev->make_current();
            settle(dev);
            CUDA_CHECK(cudaMemcpyAsync(s->window, t->dptr, t->size, cudaMemcpyDeviceToHost, dev->joined()));   // target.copy_from_slice(&*source)
            CUDA_CHECK(cudaStreamSynchronize(dev->joined()));
        }
        s->presents++;
        if (s->target.on_present)                                        // `buffer_mut().present()` (surface.rs:192)
            s->target.on_present(s->target.user_data, s->window, t->desc.width, t->desc.height, t->desc.width * t->bpp);
    });
}
wgb_status wgb_surface_texture_discard(wgb_surface surface) {
    return guarded([&] { from_handle<Surface>(surface, "surface"); });   // surface.rs:195-197: nop
}
wgb_status wgb_surface_get_window_buffer(wgb_surface surface, const void** out_pixels, uint64_t* out_size, uint64_t* out_presents) {
    return guarded([&] {
        Surface* s = from_handle<Surface>(surface, "surface");
        std::lock_guard<std::recursive_mutex> lk(s->mu);
        REQUIRE(s->configured, "Surface not configured yet");
        if (out_pixels) *out_pixels = s->window;
        if (out_size) *out_size = s->window_size;
        if (out_presents) *out_presents = s->presents;
    });
}

}  // extern "C"
