// TEST INFRASTRUCTURE (tests/cusim): host half of the software model of CUDA -- see cusim_device.h.  Implements the
// subset of the CUDA runtime / driver / NVRTC entry points that wgpu-cpu_b200/csrc/wgb_api.cpp calls, on host memory.
// "Device" allocations end at a PROT_NONE guard page, so a kernel that reads or writes past a buffer faults instead of
// passing silently.  Every stream operation executes at issue.
#include <cuda.h>
#include <cuda_runtime.h>
#include <nvrtc.h>
#undef dlopen
#undef dlsym
#include <dlfcn.h>
#undef dlopen
#undef dlsym
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#ifndef CUSIM_DIR
#error "CUSIM_DIR (absolute path of tests/cusim) must be defined"
#endif

struct cusimStream { int unused; };
struct cusimEvent { std::chrono::steady_clock::time_point t; bool recorded = false; };
struct cusimModule { void* lib; void (*launch)(void*, const char*, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, void**); };
struct cusimFunction { cusimModule* mod; void* fn; std::string name; };
struct cusimProgram { std::string source, name, log, so_path; std::vector<std::pair<std::string, std::string>> headers; };
struct CusimTexture { const void* ptr; unsigned long long bytes; };

namespace {
std::mutex g_mu;
struct Alloc { void* base; size_t len; int fd; };
std::map<void*, Alloc> g_allocs;
thread_local cudaError_t g_last = cudaSuccess;
const size_t PAGE = 4096;

// CUSIM_IPC=1: allocations are backed by a memfd, so that cudaIpcGetMemHandle / cudaIpcOpenMemHandle can map them into
// another process of the same user (the peer-memory presenter of the multi-GPU path, one process per "device")
bool ipc_enabled() { static const bool on = getenv("CUSIM_IPC") != nullptr; return on; }
int device_count() { static const int n = getenv("CUSIM_DEVICES") ? std::max(1, atoi(getenv("CUSIM_DEVICES"))) : 1; return n; }
struct IpcHandle { unsigned magic; int pid, fd; unsigned long long offset, len; };
static_assert(sizeof(IpcHandle) <= 64, "fits a cudaIpcMemHandle_t");

cudaError_t guarded_alloc(void** out, size_t size) {
    const size_t body = (size + 15) & ~(size_t)15;
    const size_t pages = (body + PAGE - 1) / PAGE + 1;
    int fd = -1;
    char* base;
    if (ipc_enabled()) {
        fd = memfd_create("cusim", 0);
        if (fd < 0 || ftruncate(fd, (off_t)(pages * PAGE)) != 0) { if (fd >= 0) close(fd); return g_last = cudaErrorMemoryAllocation; }
        base = (char*)mmap(nullptr, pages * PAGE, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    } else base = (char*)mmap(nullptr, pages * PAGE, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (base == MAP_FAILED) { if (fd >= 0) close(fd); return g_last = cudaErrorMemoryAllocation; }
    mprotect(base + (pages - 1) * PAGE, PAGE, PROT_NONE);
    char* p = base + (pages - 1) * PAGE - body;
    memset(p, 0xCD, body);                      // fresh device memory is not zero
    std::lock_guard<std::mutex> lk(g_mu);
    g_allocs[p] = Alloc{base, pages * PAGE, fd};
    *out = p;
    return cudaSuccess;
}
cudaError_t guarded_free(void* p) {
    if (!p) return cudaSuccess;
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_allocs.find(p);
    if (it == g_allocs.end()) return g_last = cudaErrorInvalidValue;
    munmap(it->second.base, it->second.len);
    if (it->second.fd >= 0) close(it->second.fd);
    g_allocs.erase(it);
    return cudaSuccess;
}
unsigned long long fnv(const std::string& s, unsigned long long h) {
    for (unsigned char c : s) { h ^= c; h *= 1099511628211ull; }
    return h;
}
std::string read_file(const std::string& path) {
    std::string out;
    if (FILE* f = fopen(path.c_str(), "rb")) { char buf[65536]; size_t n; while ((n = fread(buf, 1, sizeof buf, f)) > 0) out.append(buf, n); fclose(f); }
    return out;
}
bool write_file(const std::string& path, const std::string& text) {
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) return false;
    const bool ok = fwrite(text.data(), 1, text.size(), f) == text.size();
    fclose(f);
    return ok;
}
}  // namespace

extern "C" {
const char* cudaGetErrorString(cudaError_t e) {
    switch (e) { case cudaSuccess: return "no error"; case cudaErrorInvalidValue: return "invalid argument"; case cudaErrorMemoryAllocation: return "out of memory";
                 case cudaErrorNotSupported: return "operation not supported by the software model"; default: return "unknown error"; }
}
cudaError_t cudaGetLastError() { cudaError_t e = g_last; g_last = cudaSuccess; return e; }
static thread_local int g_device = 0;
cudaError_t cudaGetDeviceCount(int* n) { *n = device_count(); return cudaSuccess; }
cudaError_t cudaGetDevice(int* d) { *d = g_device; return cudaSuccess; }
cudaError_t cudaSetDevice(int d) { if (d < 0 || d >= device_count()) return g_last = cudaErrorInvalidValue; g_device = d; return cudaSuccess; }
cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) {
    memset(p, 0, sizeof *p);
    snprintf(p->name, sizeof p->name, "cusim software model of sm_100a");
    p->major = 10; p->minor = 0; p->multiProcessorCount = 148; p->totalGlobalMem = (size_t)8 << 30;
    return cudaSuccess;
}
cudaError_t cudaMalloc(void** p, size_t n) { return guarded_alloc(p, n ? n : 1); }
cudaError_t cudaFree(void* p) { return guarded_free(p); }
cudaError_t cudaMallocHost(void** p, size_t n) { *p = calloc(1, n ? n : 1); return *p ? cudaSuccess : (g_last = cudaErrorMemoryAllocation); }
cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { if (n) memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { if (n) memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpy2DAsync(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h, cudaMemcpyKind, cudaStream_t) {
    for (size_t r = 0; r < h; r++) memmove((char*)d + r * dp, (const char*)s + r * sp, w);
    return cudaSuccess;
}
cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { if (n) memset(d, v, n); return cudaSuccess; }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = new cusimStream(); return cudaSuccess; }
cudaError_t cudaStreamCreateWithPriority(cudaStream_t* s, unsigned, int) { *s = new cusimStream(); return cudaSuccess; }
cudaError_t cudaDeviceGetStreamPriorityRange(int* least, int* greatest) { *least = 0; *greatest = -5; return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t s) { delete s; return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new cusimEvent(); return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = new cusimEvent(); return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { e->t = std::chrono::steady_clock::now(); e->recorded = true; return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventQuery(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) {
    *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count();
    return cudaSuccess;
}
cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* a, const void* p) {
    memset(a, 0, sizeof *a);
    // CUSIM_ALL_PINNED=1: every host pointer counts as page-locked (the dry run of bench.py's pinned-upload path)
    static const bool all_pinned = getenv("CUSIM_ALL_PINNED") != nullptr;
    a->type = all_pinned ? cudaMemoryTypeHost : cudaMemoryTypeUnregistered; a->hostPointer = const_cast<void*>(p);
    return cudaSuccess;
}
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* out, void* p) {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_allocs.find(p);
    if (it == g_allocs.end()) return g_last = cudaErrorInvalidValue;
    if (it->second.fd < 0) return g_last = cudaErrorNotSupported;             // needs CUSIM_IPC=1
    IpcHandle h{0x43534950u, (int)getpid(), it->second.fd, (unsigned long long)((char*)p - (char*)it->second.base), (unsigned long long)it->second.len};
    memset(out, 0, sizeof *out);
    memcpy(out, &h, sizeof h);
    return cudaSuccess;
}
cudaError_t cudaIpcOpenMemHandle(void** out, cudaIpcMemHandle_t handle, unsigned) {
    IpcHandle h;
    memcpy(&h, &handle, sizeof h);
    if (h.magic != 0x43534950u) return g_last = cudaErrorInvalidValue;
    char path[64];
    snprintf(path, sizeof path, "/proc/%d/fd/%d", h.pid, h.fd);
    const int fd = open(path, O_RDWR);
    if (fd < 0) return g_last = cudaErrorInvalidValue;
    char* base = (char*)mmap(nullptr, h.len, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (base == MAP_FAILED) return g_last = cudaErrorMemoryAllocation;
    mprotect(base + h.len - PAGE, PAGE, PROT_NONE);
    std::lock_guard<std::mutex> lk(g_mu);
    g_allocs[base + h.offset] = Alloc{base, (size_t)h.len, -1};
    *out = base + h.offset;
    return cudaSuccess;
}
cudaError_t cudaIpcCloseMemHandle(void* p) { return guarded_free(p); }
cudaError_t cudaCreateTextureObject(cudaTextureObject_t* out, const cudaResourceDesc* rd, const cudaTextureDesc*, const void*) {
    if (rd->resType != cudaResourceTypeLinear) return g_last = cudaErrorNotSupported;
    *out = (cudaTextureObject_t)(uintptr_t) new CusimTexture{rd->res.linear.devPtr, rd->res.linear.sizeInBytes};
    return cudaSuccess;
}
cudaError_t cudaDestroyTextureObject(cudaTextureObject_t t) { delete (CusimTexture*)(uintptr_t)t; return cudaSuccess; }

// ---- driver -----------------------------------------------------------------------------------------------
CUresult cusim_cuModuleLoadData(CUmodule* out, const void* image) {
    const char* path = (const char*)image;                       // the stand-in "cubin" is the path of the shared object
    void* lib = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!lib) { fprintf(stderr, "cusim: dlopen(%s): %s\n", path, dlerror()); return CUDA_ERROR_INVALID_VALUE; }
    auto launch = (decltype(cusimModule::launch))dlsym(lib, "cusim_launch");
    if (!launch) { dlclose(lib); return CUDA_ERROR_NOT_FOUND; }
    *out = new cusimModule{lib, launch};
    return CUDA_SUCCESS;
}
CUresult cusim_cuModuleUnload(CUmodule m) { if (m) { dlclose(m->lib); delete m; } return CUDA_SUCCESS; }
CUresult cusim_cuModuleGetFunction(CUfunction* out, CUmodule m, const char* name) {
    void* fn = dlsym(m->lib, name);
    if (!fn) return CUDA_ERROR_NOT_FOUND;
    *out = new cusimFunction{m, fn, name};                       // lives as long as the process (a handful per pipeline)
    return CUDA_SUCCESS;
}
CUresult cusim_cuLaunchKernel(CUfunction f, unsigned gx, unsigned gy, unsigned gz, unsigned bx, unsigned by, unsigned bz, unsigned, CUstream, void** params, void**) {
    if (!f || !gx || !gy || !gz || !bx || !by || !bz || (unsigned long long)bx * by * bz > 1024ull) return CUDA_ERROR_INVALID_VALUE;
    f->mod->launch(f->fn, f->name.c_str(), gx, gy, gz, bx, by, bz, params);
    return CUDA_SUCCESS;
}
CUresult cusim_cuGetErrorString(CUresult r, const char** s) {
    *s = r == CUDA_SUCCESS ? "no error" : r == CUDA_ERROR_NOT_FOUND ? "named symbol not found" : r == CUDA_ERROR_INVALID_VALUE ? "invalid value" : "unknown driver error";
    return CUDA_SUCCESS;
}
// the alignment and extent rules of the real call (a map the hardware would reject must not pass here either)
CUresult cusim_cuTensorMapEncodeTiled(CUtensorMap* out, CUtensorMapDataType type, cuuint32_t rank, void* base, const cuuint64_t* dims, const cuuint64_t* strides,
                                      const cuuint32_t* box, const cuuint32_t* elem, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill) {
    if (!out || type != CU_TENSOR_MAP_DATA_TYPE_UINT32 || rank != 2 || !base || ((uintptr_t)base & 15u) || (strides[0] & 15u) || !dims[0] || !dims[1] ||
        box[0] == 0 || box[0] > 256 || box[1] == 0 || box[1] > 256 || (box[0] * 4u) % 16u || elem[0] != 1 || elem[1] != 1 || strides[0] < dims[0] * 4)
        return CUDA_ERROR_INVALID_VALUE;
    memset(out, 0, sizeof *out);
    out->opaque[0] = (cuuint64_t)(uintptr_t)base; out->opaque[1] = dims[0]; out->opaque[2] = dims[1]; out->opaque[3] = strides[0];
    out->opaque[4] = box[0]; out->opaque[5] = box[1];
    return CUDA_SUCCESS;
}
cudaError_t cudaGetDriverEntryPoint(const char* name, void** fn, unsigned long long, cudaDriverEntryPointQueryResult* q) {
    *fn = nullptr;
    if (!strcmp(name, "cuModuleLoadData")) *fn = (void*)&cusim_cuModuleLoadData;
    else if (!strcmp(name, "cuModuleUnload")) *fn = (void*)&cusim_cuModuleUnload;
    else if (!strcmp(name, "cuModuleGetFunction")) *fn = (void*)&cusim_cuModuleGetFunction;
    else if (!strcmp(name, "cuLaunchKernel")) *fn = (void*)&cusim_cuLaunchKernel;
    else if (!strcmp(name, "cuGetErrorString")) *fn = (void*)&cusim_cuGetErrorString;
    else if (!strcmp(name, "cuTensorMapEncodeTiled")) *fn = (void*)&cusim_cuTensorMapEncodeTiled;
    if (q) *q = *fn ? cudaDriverEntryPointSuccess : cudaDriverEntryPointSymbolNotFound;
    return cudaSuccess;
}

// ---- "NVRTC": g++ over the pipeline TU with the device half of the model force-included ---------------------------
nvrtcResult nvrtcCreateProgram(nvrtcProgram* out, const char* src, const char* name, int nh, const char* const* headers, const char* const* names) {
    auto* p = new cusimProgram();
    p->source = src; p->name = name ? name : "program.cu";
    for (int i = 0; i < nh; i++) p->headers.emplace_back(names[i], headers[i]);
    *out = p;
    return NVRTC_SUCCESS;
}
nvrtcResult nvrtcDestroyProgram(nvrtcProgram* p) { delete *p; *p = nullptr; return NVRTC_SUCCESS; }
nvrtcResult nvrtcCompileProgram(nvrtcProgram p, int, const char* const*) {
    const std::string dev = read_file(std::string(CUSIM_DIR) + "/cusim_device.h"), entry = read_file(std::string(CUSIM_DIR) + "/cusim_entry.h");
    if (dev.empty() || entry.empty()) { p->log = "cusim: cannot read cusim_device.h / cusim_entry.h under " CUSIM_DIR; return NVRTC_ERROR_COMPILATION; }
    const char* opt = getenv("CUSIM_OPT");
    const std::string flags = std::string("-std=c++17 -fPIC -shared -g1 -w -ffp-contract=off -fno-strict-aliasing -fno-math-errno ") + (opt ? opt : "-O1");
    unsigned long long h1 = fnv(p->source, 14695981039346656037ull), h2 = fnv(p->source, 0x9E3779B97F4A7C15ull);
    for (auto& kv : p->headers) { h1 = fnv(kv.first, fnv(kv.second, h1)); h2 = fnv(kv.second, fnv(kv.first, h2)); }
    h1 = fnv(flags, fnv(entry, fnv(dev, h1))); h2 = fnv(dev, fnv(entry, fnv(flags, h2)));
    const char* cache_env = getenv("CUSIM_CACHE");
    const std::string cache = cache_env ? cache_env : "/tmp/cusim_cache_" + std::to_string((unsigned)getuid());
    mkdir(cache.c_str(), 0700);
    char key[40];
    snprintf(key, sizeof key, "%016llx%016llx", h1, h2);
    const std::string so = cache + "/" + key + ".so";
    if (access(so.c_str(), R_OK) != 0) {
        static std::atomic<unsigned> serial{0};
        const std::string dir = cache + "/" + key + "." + std::to_string((long)getpid()) + "." + std::to_string(serial.fetch_add(1));
        mkdir(dir.c_str(), 0700);
        bool ok = write_file(dir + "/cusim_device.h", dev) && write_file(dir + "/cusim_entry.h", entry) && write_file(dir + "/" + p->name, p->source);
        for (auto& kv : p->headers) ok = ok && write_file(dir + "/" + kv.first, kv.second);
        ok = ok && write_file(dir + "/tu.cpp", "#include \"cusim_device.h\"\n#include \"" + p->name + "\"\n#include \"cusim_entry.h\"\n");
        if (!ok) { p->log = "cusim: cannot write " + dir; return NVRTC_ERROR_COMPILATION; }
        const std::string cmd = "cd '" + dir + "' && g++ " + flags + " -I. tu.cpp -o out.so -lpthread > log.txt 2>&1";
        const int rc = system(cmd.c_str());
        p->log = read_file(dir + "/log.txt");
        if (rc != 0) { if (p->log.size() > 20000) p->log.resize(20000); return NVRTC_ERROR_COMPILATION; }
        if (rename((dir + "/out.so").c_str(), so.c_str()) != 0) { p->log = "cusim: rename failed"; return NVRTC_ERROR_COMPILATION; }
        if (!getenv("CUSIM_KEEP")) { const std::string rm = "rm -rf '" + dir + "'"; if (system(rm.c_str())) {} }
    }
    p->so_path = so;
    return NVRTC_SUCCESS;
}
nvrtcResult nvrtcGetCUBINSize(nvrtcProgram p, size_t* n) { *n = p->so_path.size() + 1; return NVRTC_SUCCESS; }
nvrtcResult nvrtcGetCUBIN(nvrtcProgram p, char* out) { memcpy(out, p->so_path.c_str(), p->so_path.size() + 1); return NVRTC_SUCCESS; }
nvrtcResult nvrtcGetProgramLogSize(nvrtcProgram p, size_t* n) { *n = p->log.size() + 1; return NVRTC_SUCCESS; }
nvrtcResult nvrtcGetProgramLog(nvrtcProgram p, char* out) { memcpy(out, p->log.c_str(), p->log.size() + 1); return NVRTC_SUCCESS; }
const char* nvrtcGetErrorString(nvrtcResult r) { return r == NVRTC_SUCCESS ? "NVRTC_SUCCESS" : "NVRTC_ERROR_COMPILATION"; }

// ---- the two dlfcn calls wgb_api.cpp makes to find NVRTC ---------------------------------------------------------
static int g_nvrtc_handle;
void* cusim_dlopen(const char* name, int flags) {
    if (name && strstr(name, "nvrtc")) return &g_nvrtc_handle;
    return dlopen(name, flags);
}
void* cusim_dlsym(void* lib, const char* name) {
    if (lib != &g_nvrtc_handle) return dlsym(lib, name);
#define S(n) if (!strcmp(name, #n)) return (void*)&n;
    S(nvrtcCreateProgram) S(nvrtcDestroyProgram) S(nvrtcCompileProgram) S(nvrtcGetCUBINSize) S(nvrtcGetCUBIN)
    S(nvrtcGetProgramLogSize) S(nvrtcGetProgramLog) S(nvrtcGetErrorString)
#undef S
    return nullptr;
}
}  // extern "C"
