// cusim_device.h -- TEST INFRASTRUCTURE: a software model of the CUDA execution model, used to run this repository's
// kernels (wgb_raster.cuh / wgb_prelude.cuh, unchanged) on the host cores of a machine that has no GPU, so that the
// kernel source itself -- not a restatement of it -- is checked against the oracle by the CPU test tier.
//
// It is NOT a product path: nothing under wgpu-cpu_b200/ includes or loads it, the shipped libwgpu_b200.so is built
// against the real CUDA runtime and fails loudly without a device.  tests/cusim/build.py compiles the same host
// runtime source against tests/cusim/include/ (stand-ins for cuda_runtime.h / cuda.h / nvrtc.h) into
// tests/cusim/_build/libwgpu_b200_sim.so; its "NVRTC" compiles each pipeline translation unit with g++ after
// force-including this header, and its "cuLaunchKernel" runs the grid below.
//
// Model: a CTA is a set of fibers (one per thread) on one OS thread; CTAs are spread over a pool of OS threads.
// A fiber runs until it blocks in __syncthreads or a *_sync warp collective (or finishes); the scheduler then
// resumes the next runnable fiber.  A full pass over the CTA in which no fiber can run is reported as a deadlock
// (a barrier not reached by every live thread, a collective whose mask names a thread that never arrives).
// Float arithmetic: the TU is compiled with -ffp-contract=off, so +,-,*,/ are single IEEE binary32 operations as
// with nvcc --fmad=false; __fdividef / __frcp_rn / ex2 etc. are modelled by the correctly rounded operation.
#pragma once
#define WGB_CUSIM 1

#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>

#include <algorithm>
#include <atomic>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __grid_constant__
#define __restrict__ __restrict
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static thread_local
#define __constant__ static const

struct uint3 { unsigned x, y, z; };
struct dim3 { unsigned x, y, z; };
struct float2 { float x, y; };
struct __attribute__((aligned(16))) float4 { float x, y, z, w; };
struct uint2 { unsigned x, y; };
struct __attribute__((aligned(16))) uint4 { unsigned x, y, z, w; };
struct uchar4 { unsigned char x, y, z, w; };
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
static inline uchar4 make_uchar4(unsigned char x, unsigned char y, unsigned char z, unsigned char w) { return uchar4{x, y, z, w}; }
typedef unsigned long long cudaTextureObject_t;

// ------------------------------------------------------------------------------------------------------------
// fibers
// ------------------------------------------------------------------------------------------------------------
extern "C" void cusim_switch(void** save_sp, void* load_sp) __attribute__((visibility("hidden")));
__asm__(
    ".text\n"
    ".type cusim_switch,@function\n"
    "cusim_switch:\n"
    "  pushq %rbp\n  pushq %rbx\n  pushq %r12\n  pushq %r13\n  pushq %r14\n  pushq %r15\n"
    "  movq %rsp, (%rdi)\n"
    "  movq %rsi, %rsp\n"
    "  popq %r15\n  popq %r14\n  popq %r13\n  popq %r12\n  popq %rbx\n  popq %rbp\n"
    "  ret\n"
    ".size cusim_switch,.-cusim_switch\n");

enum { CUSIM_RUN = 0, CUSIM_WAIT_BARRIER = 1, CUSIM_WAIT_WARP = 2, CUSIM_DONE = 3, CUSIM_POLL = 4 };
enum { CUSIM_OP_SHFL = 0, CUSIM_OP_SHFL_XOR, CUSIM_OP_SHFL_UP, CUSIM_OP_SHFL_DOWN, CUSIM_OP_BALLOT, CUSIM_OP_MATCH_ANY,
       CUSIM_OP_SYNCWARP, CUSIM_OP_ANY, CUSIM_OP_ALL };

struct CusimWarp {
    unsigned arrived, exited, done;
    unsigned long long val[32], res[32];
    unsigned mask[32];
    int op[32], arg[32], width[32];
};
struct CusimLane {
    void* sp;
    uint3 tid;
    unsigned lane, warp, linear;
    int state;
    unsigned barrier_gen;
    char* stack;
};
struct CusimBlock {
    uint3 bid;
    dim3 bdim, gdim;
    unsigned nthreads, alive, barrier_arrived, barrier_gen;
    void* sched_sp;
    CusimLane* lanes;
    CusimWarp* warps;
    const void* body;         // std::function-like: called through cusim_body_call
    void (*body_call)(const void*);
    unsigned long long switches;
};
static thread_local CusimBlock* cusim_blk;
static thread_local CusimLane* cusim_cur;

#define threadIdx (cusim_cur->tid)
#define blockIdx (cusim_blk->bid)
#define blockDim (cusim_blk->bdim)
#define gridDim (cusim_blk->gdim)
#define warpSize 32

static const size_t CUSIM_STACK = 96 * 1024;

static __attribute__((noinline)) void cusim_yield() {
    CusimLane* me = cusim_cur;
    cusim_blk->switches++;
    cusim_switch(&me->sp, cusim_blk->sched_sp);
}

static void cusim_warp_try_complete(CusimWarp& w, unsigned mask) {
    if (((w.arrived | w.exited) & mask) != mask) return;
    const unsigned part = w.arrived & mask;
    if (!part) return;
    unsigned ballot = 0;
    for (unsigned i = 0; i < 32; i++)
        if (part >> i & 1u) {
            if (w.mask[i] != mask) {
                fprintf(stderr, "cusim: lanes of one warp meet in a collective with different masks (%08x vs %08x)\n", w.mask[i], mask);
                abort();
            }
            if (w.val[i]) ballot |= 1u << i;
        }
    for (unsigned i = 0; i < 32; i++) {
        if (!(part >> i & 1u)) continue;
        const int width = w.width[i] ? w.width[i] : 32;
        const unsigned seg = i & ~(unsigned)(width - 1);
        unsigned long long r = w.val[i];
        switch (w.op[i]) {
        case CUSIM_OP_SHFL: { const unsigned s = seg | ((unsigned)w.arg[i] & (unsigned)(width - 1)); if (part >> s & 1u) r = w.val[s]; break; }
        case CUSIM_OP_SHFL_XOR: { const unsigned s = i ^ (unsigned)w.arg[i]; if (s < 32 && (s & ~(unsigned)(width - 1)) == seg && (part >> s & 1u)) r = w.val[s]; break; }
        case CUSIM_OP_SHFL_UP: { const int s = (int)i - w.arg[i]; if (s >= (int)seg && (part >> s & 1u)) r = w.val[s]; break; }
        case CUSIM_OP_SHFL_DOWN: { const unsigned s = i + (unsigned)w.arg[i]; if (s < seg + (unsigned)width && (part >> s & 1u)) r = w.val[s]; break; }
        case CUSIM_OP_BALLOT: r = ballot; break;
        case CUSIM_OP_ANY: r = ballot != 0; break;
        case CUSIM_OP_ALL: r = ballot == part; break;
        case CUSIM_OP_MATCH_ANY: { unsigned m = 0; for (unsigned j = 0; j < 32; j++) if ((part >> j & 1u) && w.val[j] == w.val[i]) m |= 1u << j; r = m; break; }
        default: break;
        }
        w.res[i] = r;
    }
    w.done |= part;
    w.arrived &= ~part;
}

static unsigned long long cusim_collective(int op, unsigned mask, unsigned long long val, int arg, int width) {
    CusimLane* me = cusim_cur;
    CusimWarp& w = cusim_blk->warps[me->warp];
    const unsigned bit = 1u << me->lane;
    if (!(mask & bit)) { fprintf(stderr, "cusim: a lane calls a *_sync collective whose mask does not name it\n"); abort(); }
    w.val[me->lane] = val; w.mask[me->lane] = mask; w.op[me->lane] = op; w.arg[me->lane] = arg; w.width[me->lane] = width;
    w.arrived |= bit;
    cusim_warp_try_complete(w, mask);
    while (!(w.done & bit)) { me->state = CUSIM_WAIT_WARP; cusim_yield(); }
    me->state = CUSIM_RUN;
    w.done &= ~bit;
    return w.res[me->lane];
}

static void cusim_barrier_release(CusimBlock* b) { b->barrier_arrived = 0; b->barrier_gen++; }

static inline void __syncthreads() {
    CusimBlock* b = cusim_blk;
    CusimLane* me = cusim_cur;
    const unsigned gen = b->barrier_gen;
    if (++b->barrier_arrived == b->alive) { cusim_barrier_release(b); return; }
    me->barrier_gen = gen;
    while (b->barrier_gen == gen) { me->state = CUSIM_WAIT_BARRIER; cusim_yield(); }
    me->state = CUSIM_RUN;
}

static void cusim_lane_exit() {
    CusimBlock* b = cusim_blk;
    CusimLane* me = cusim_cur;
    me->state = CUSIM_DONE;
    b->alive--;
    if (b->alive && b->barrier_arrived == b->alive) cusim_barrier_release(b);
    CusimWarp& w = b->warps[me->warp];
    w.exited |= 1u << me->lane;
    for (unsigned i = 0; i < 32; i++)
        if (w.arrived >> i & 1u) cusim_warp_try_complete(w, w.mask[i]);
    void* dummy;
    cusim_switch(&dummy, b->sched_sp);
    abort();
}

static void cusim_lane_entry() {
    cusim_blk->body_call(cusim_blk->body);
    cusim_lane_exit();
}

struct CusimWorker {
    char* stacks = nullptr;
    size_t nstacks = 0;
    std::vector<CusimLane> lanes;
    std::vector<CusimWarp> warps;
    ~CusimWorker() { if (stacks) munmap(stacks, nstacks * CUSIM_STACK); }
    void reserve(unsigned n) {
        if (n > nstacks) {
            if (stacks) munmap(stacks, nstacks * CUSIM_STACK);
            stacks = (char*)mmap(nullptr, (size_t)n * CUSIM_STACK, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
            if (stacks == MAP_FAILED) { perror("cusim: mmap"); abort(); }
            nstacks = n;
        }
        lanes.resize(n);
        warps.resize((n + 31) / 32);
    }
};

// runs one CTA to completion on the calling OS thread
static void cusim_run_block(CusimWorker& wk, dim3 grid, dim3 block, uint3 bid, const void* body, void (*body_call)(const void*)) {
    const unsigned n = block.x * block.y * block.z;
    wk.reserve(n);
    CusimBlock b;
    memset(&b, 0, sizeof b);
    b.bid = bid; b.bdim = block; b.gdim = grid; b.nthreads = n; b.alive = n;
    b.lanes = wk.lanes.data(); b.warps = wk.warps.data(); b.body = body; b.body_call = body_call;
    memset(b.warps, 0, sizeof(CusimWarp) * wk.warps.size());
    if (n & 31u) b.warps[n / 32].exited = ~0u << (n & 31u);         // lanes past the end of a partial last warp
    for (unsigned t = 0; t < n; t++) {
        CusimLane& l = b.lanes[t];
        l.linear = t; l.lane = t & 31u; l.warp = t >> 5; l.state = CUSIM_RUN; l.barrier_gen = 0;
        l.tid.x = t % block.x; l.tid.y = (t / block.x) % block.y; l.tid.z = t / (block.x * block.y);
        l.stack = wk.stacks + (size_t)t * CUSIM_STACK;
        void** top = (void**)(l.stack + CUSIM_STACK);
        top[-1] = nullptr;                          // fake return address: keeps rsp % 16 == 8 at the entry
        top[-2] = (void*)&cusim_lane_entry;
        for (int k = 3; k <= 8; k++) top[-k] = nullptr;
        l.sp = (void*)(top - 8);
    }
    cusim_blk = &b;
    unsigned done = 0;
    // CUSIM_SCHED_SEED != 0 shuffles the schedule: every pass starts at a random thread, walks up or down, and leaves
    // out a random quarter of the runnable threads -- the kernels must give the same bytes for every legal
    // interleaving (missing barriers and lane-order assumptions show up as a parity failure under some seed)
    static const unsigned long long sched_seed = getenv("CUSIM_SCHED_SEED") ? strtoull(getenv("CUSIM_SCHED_SEED"), nullptr, 0) : 0ull;
    unsigned long long rng = sched_seed * 0x9E3779B97F4A7C15ull + (((unsigned long long)bid.x << 32) ^ ((unsigned long long)bid.y << 16) ^ bid.z ^ 0xD1B54A32D192ED03ull);
    auto next_rand = [&]() { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; return rng; };
    while (done < n) {
        bool progressed = false;
        unsigned start = 0; bool down = false;
        if (sched_seed) { const unsigned long long r = next_rand(); start = (unsigned)(r % n); down = (r >> 40) & 1u; }
        for (unsigned k = 0; k < n; k++) {
            const unsigned t = sched_seed ? (down ? (start + n - k) % n : (start + k) % n) : k;
            CusimLane& l = b.lanes[t];
            if (l.state == CUSIM_DONE) continue;
            if (sched_seed && progressed && (next_rand() & 3u) == 0u) continue;
            if (l.state == CUSIM_WAIT_BARRIER && b.barrier_gen == l.barrier_gen) continue;
            if (l.state == CUSIM_WAIT_WARP && !(b.warps[l.warp].done >> l.lane & 1u)) continue;
            const int before = l.state;
            cusim_cur = &l;
            cusim_switch(&b.sched_sp, l.sp);
            if (!(before == CUSIM_POLL && l.state == CUSIM_POLL)) progressed = true;    // a lane that polls and polls again made no progress
            if (l.state == CUSIM_DONE) done++;
        }
        if (!progressed) {
            fprintf(stderr, "cusim: deadlock in block (%u,%u,%u): %u of %u threads alive, %u at the barrier\n", bid.x, bid.y, bid.z, b.alive, n, b.barrier_arrived);
            for (unsigned t = 0; t < n && t < 64; t++)
                if (b.lanes[t].state != CUSIM_DONE) fprintf(stderr, "  thread %u state %d\n", t, b.lanes[t].state);
            abort();
        }
    }
    cusim_blk = nullptr; cusim_cur = nullptr;
}

template <class F>
static void cusim_body_thunk(const void* f) { (*(const F*)f)(); }

template <class F>
static void cusim_run_grid(dim3 grid, dim3 block, const F& body) {
    const unsigned long long total = (unsigned long long)grid.x * grid.y * grid.z;
    unsigned nw = std::thread::hardware_concurrency();
    if (const char* e = getenv("CUSIM_THREADS")) nw = (unsigned)atoi(e);
    if (nw < 1) nw = 1;
    if (nw > total) nw = (unsigned)total;
    std::atomic<unsigned long long> next{0};
    auto work = [&]() {
        CusimWorker wk;
        for (;;) {
            const unsigned long long i = next.fetch_add(1);
            if (i >= total) break;
            uint3 bid{(unsigned)(i % grid.x), (unsigned)((i / grid.x) % grid.y), (unsigned)(i / ((unsigned long long)grid.x * grid.y))};
            cusim_run_block(wk, grid, block, bid, &body, &cusim_body_thunk<F>);
        }
    };
    if (nw == 1) { work(); return; }
    std::vector<std::thread> th;
    for (unsigned i = 0; i < nw; i++) th.emplace_back(work);
    for (auto& t : th) t.join();
}

// ------------------------------------------------------------------------------------------------------------
// warp collectives
// ------------------------------------------------------------------------------------------------------------
template <class T> static inline unsigned long long cusim_bits(T v) { unsigned long long b = 0; static_assert(sizeof(T) <= 8, ""); memcpy(&b, &v, sizeof(T)); return b; }
template <class T> static inline T cusim_unbits(unsigned long long b) { T v; memcpy(&v, &b, sizeof(T)); return v; }
template <class T> static inline T __shfl_sync(unsigned m, T v, int src, int width = 32) { return cusim_unbits<T>(cusim_collective(CUSIM_OP_SHFL, m, cusim_bits(v), src, width)); }
template <class T> static inline T __shfl_xor_sync(unsigned m, T v, int lm, int width = 32) { return cusim_unbits<T>(cusim_collective(CUSIM_OP_SHFL_XOR, m, cusim_bits(v), lm, width)); }
template <class T> static inline T __shfl_up_sync(unsigned m, T v, unsigned d, int width = 32) { return cusim_unbits<T>(cusim_collective(CUSIM_OP_SHFL_UP, m, cusim_bits(v), (int)d, width)); }
template <class T> static inline T __shfl_down_sync(unsigned m, T v, unsigned d, int width = 32) { return cusim_unbits<T>(cusim_collective(CUSIM_OP_SHFL_DOWN, m, cusim_bits(v), (int)d, width)); }
static inline unsigned __ballot_sync(unsigned m, int pred) { return (unsigned)cusim_collective(CUSIM_OP_BALLOT, m, pred != 0, 0, 32); }
static inline int __any_sync(unsigned m, int pred) { return (int)cusim_collective(CUSIM_OP_ANY, m, pred != 0, 0, 32); }
static inline int __all_sync(unsigned m, int pred) { return (int)cusim_collective(CUSIM_OP_ALL, m, pred != 0, 0, 32); }
template <class T> static inline unsigned __match_any_sync(unsigned m, T v) { return (unsigned)cusim_collective(CUSIM_OP_MATCH_ANY, m, cusim_bits(v), 0, 32); }
static inline void __syncwarp(unsigned m = 0xFFFFFFFFu) {
    // exited lanes and lanes past the block end are not waited for (as on the hardware)
    cusim_collective(CUSIM_OP_SYNCWARP, m, 0, 0, 32);
}
// Which lanes are converged is a property of the hardware scheduler.  The model answers "this lane alone", the one
// answer that is always valid: code built on __activemask() must be correct for every grouping of the lanes.
static inline unsigned __activemask() { return 1u << cusim_cur->lane; }
static inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline void __threadfence_block() {}

// ------------------------------------------------------------------------------------------------------------
// atomics (CTAs run on different OS threads: real atomics)
// ------------------------------------------------------------------------------------------------------------
template <class T> static inline T atomicAdd(T* p, T v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline float atomicAdd(float* p, float v) {
    unsigned* u = (unsigned*)p; unsigned o = __atomic_load_n(u, __ATOMIC_RELAXED), n;
    do { float f = cusim_unbits<float>(o) + v; n = (unsigned)cusim_bits(f); } while (!__atomic_compare_exchange_n(u, &o, n, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
    return cusim_unbits<float>(o);
}
template <class T> static inline T atomicOr(T* p, T v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
template <class T> static inline T atomicAnd(T* p, T v) { return __atomic_fetch_and(p, v, __ATOMIC_RELAXED); }
template <class T> static inline T atomicExch(T* p, T v) { return __atomic_exchange_n(p, v, __ATOMIC_RELAXED); }
template <class T> static inline T atomicCAS(T* p, T c, T v) { __atomic_compare_exchange_n(p, &c, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED); return c; }
template <class T> static inline T atomicMin(T* p, T v) {
    T o = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (v < o && !__atomic_compare_exchange_n(p, &o, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return o;
}
template <class T> static inline T atomicMax(T* p, T v) {
    T o = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (v > o && !__atomic_compare_exchange_n(p, &o, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return o;
}

// ------------------------------------------------------------------------------------------------------------
// intrinsics
// ------------------------------------------------------------------------------------------------------------
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fdividef(float a, float b) { return a / b; }
static inline float __frcp_rn(float a) { return 1.0f / a; }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline float __int2float_rn(int a) { return (float)a; }
static inline float __uint2float_rn(unsigned a) { return (float)a; }
static inline float __ll2float_rn(long long a) { return (float)a; }
static inline float __ull2float_rn(unsigned long long a) { return (float)a; }
static inline double __ll2double_rn(long long a) { return (double)a; }
static inline int __float2int_rz(float v) {
    if (v != v) return 0;
    if (v >= 2147483648.0f) return 2147483647;
    if (v <= -2147483648.0f) return (int)0x80000000;
    return (int)v;
}
static inline unsigned __float2uint_rz(float v) {
    if (!(v > 0.0f)) return 0u;          // NaN, negatives and zero
    if (v >= 4294967296.0f) return 0xFFFFFFFFu;
    return (unsigned)v;
}
static inline int __float2int_rn(float v) { return __float2int_rz(nearbyintf(v)); }
static inline unsigned __float_as_uint(float f) { return (unsigned)cusim_bits(f); }
static inline int __float_as_int(float f) { return (int)cusim_bits(f); }
static inline float __uint_as_float(unsigned u) { return cusim_unbits<float>(u); }
static inline float __int_as_float(int u) { return cusim_unbits<float>((unsigned)u); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline unsigned __brev(unsigned v) { unsigned r = 0; for (int i = 0; i < 32; i++) r |= (v >> i & 1u) << (31 - i); return r; }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
static inline int __mulhi(int a, int b) { return (int)(((long long)a * b) >> 32); }
static inline float __saturatef(float v) { return v != v ? 0.0f : v < 0.0f ? 0.0f : v > 1.0f ? 1.0f : v; }
static inline float rsqrtf(float a) { return 1.0f / sqrtf(a); }
static inline float exp10f_(float a) { return powf(10.0f, a); }
template <class T> static inline T __ldg(const T* p) { return *p; }

// CUDA's global min / max overload set (mixed signedness promotes to unsigned)
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline unsigned min(int a, unsigned b) { return min((unsigned)a, b); }
static inline unsigned min(unsigned a, int b) { return min(a, (unsigned)b); }
static inline unsigned max(int a, unsigned b) { return max((unsigned)a, b); }
static inline unsigned max(unsigned a, int b) { return max(a, (unsigned)b); }
static inline long long min(long long a, long long b) { return a < b ? a : b; }
static inline long long max(long long a, long long b) { return a > b ? a : b; }
static inline unsigned long long min(unsigned long long a, unsigned long long b) { return a < b ? a : b; }
static inline unsigned long long max(unsigned long long a, unsigned long long b) { return a > b ? a : b; }
static inline float min(float a, float b) { return fminf(a, b); }
static inline float max(float a, float b) { return fmaxf(a, b); }

// linear textures bound by cudaCreateTextureObject (tests/cusim/cusim_host.cpp): element fetch, clamped index
struct CusimTexture { const void* ptr; unsigned long long bytes; };
template <class T> static inline T tex1Dfetch(cudaTextureObject_t obj, int i) {
    const CusimTexture* t = (const CusimTexture*)(uintptr_t)obj;
    const long long n = (long long)(t->bytes / sizeof(T));
    if (i < 0 || i >= n) { T z; memset(&z, 0, sizeof z); return z; }     // out-of-range linear fetches return zero
    return ((const T*)t->ptr)[i];
}

// TMA bulk copies and the mbarrier they complete on (wgb_raster.cuh "TMA bulk copies").  The model copies at issue;
// the mbarrier keeps the real protocol (arrival count + transaction bytes, phase parity), so a wait that is not
// covered by an expect_tx, or bytes that never arrive, show up as a deadlock here too.
#define WGB_BULK_COPY_PRIMITIVES_PROVIDED 1
struct CusimMbar { int tx; short pending; unsigned char phase, count; };
static_assert(sizeof(CusimMbar) == 8, "an mbarrier is one 64-bit shared-memory word");
static inline void cusim_mbar_check(CusimMbar* m) { if (m->pending == 0 && m->tx == 0) { m->phase ^= 1u; m->pending = m->count; } }
static inline void wgb_bulk_store_row(unsigned long long global_dst, const void* smem_src, unsigned bytes) { memcpy((void*)(uintptr_t)global_dst, smem_src, bytes); }
static inline void wgb_bulk_store_commit_and_wait() {}
static inline void wgb_bulk_load_row(void* smem_dst, unsigned long long global_src, unsigned bytes, unsigned long long* mbar) {
    memcpy(smem_dst, (const void*)(uintptr_t)global_src, bytes);
    CusimMbar* m = (CusimMbar*)mbar; m->tx -= (int)bytes; cusim_mbar_check(m);
}
static inline void wgb_mbar_init(unsigned long long* mbar, unsigned count) { CusimMbar* m = (CusimMbar*)mbar; m->tx = 0; m->pending = (short)count; m->phase = 0; m->count = (unsigned char)count; }
static inline void wgb_mbar_expect_tx(unsigned long long* mbar, unsigned bytes) { CusimMbar* m = (CusimMbar*)mbar; m->tx += (int)bytes; m->pending--; cusim_mbar_check(m); }
static inline void wgb_mbar_wait(unsigned long long* mbar, unsigned parity) {
    CusimMbar* m = (CusimMbar*)mbar;
    CusimLane* me = cusim_cur;
    while (m->phase == parity) { me->state = CUSIM_POLL; cusim_yield(); }
    me->state = CUSIM_RUN;
}
static inline void wgb_fence_proxy_async() {}
// tensor-map copies (see cusim_cuTensorMapEncodeTiled for the words): a store skips the texels of the box that lie
// outside the tensor, a load fills them with zeros and completes the full box's bytes on the mbarrier
struct WgbTensorMap;
static inline void wgb_tensor_store_tile(const WgbTensorMap* map, unsigned x, unsigned y, const void* smem_src) {
    const unsigned long long* w = (const unsigned long long*)map;
    const unsigned bw = (unsigned)w[4], bh = (unsigned)w[5];
    for (unsigned r = 0; r < bh; r++)
        for (unsigned c = 0; c < bw; c++)
            if (x + c < w[1] && y + r < w[2]) *(unsigned*)(uintptr_t)(w[0] + (unsigned long long)(y + r) * w[3] + (x + c) * 4ull) = ((const unsigned*)smem_src)[r * bw + c];
}
static inline void wgb_tensor_load_tile(void* smem_dst, const WgbTensorMap* map, unsigned x, unsigned y, unsigned long long* mbar) {
    const unsigned long long* w = (const unsigned long long*)map;
    const unsigned bw = (unsigned)w[4], bh = (unsigned)w[5];
    for (unsigned r = 0; r < bh; r++)
        for (unsigned c = 0; c < bw; c++)
            ((unsigned*)smem_dst)[r * bw + c] = (x + c < w[1] && y + r < w[2]) ? *(const unsigned*)(uintptr_t)(w[0] + (unsigned long long)(y + r) * w[3] + (x + c) * 4ull) : 0u;
    CusimMbar* m = (CusimMbar*)mbar; m->tx -= (int)(bw * bh * 4u); cusim_mbar_check(m);
}

// binary16 <-> binary32 (cvt.rn.f16.f32 / cvt.f32.f16) through the host compiler's _Float16 (round to nearest even)
#define WGB_F16_CONVERSIONS_PROVIDED 1
static inline unsigned short wgb_f32_to_f16_bits(float v) { _Float16 h = (_Float16)v; unsigned short b; memcpy(&b, &h, 2); return b; }
static inline float wgb_f16_bits_to_f32(unsigned short b) { _Float16 h; memcpy(&h, &b, 2); return (float)h; }

// MUFU.RCP (rcp.approx.ftz.f32, at most 1 ulp off) is modelled by the correctly rounded reciprocal
#define WGB_RCP_APPROX_PROVIDED 1
static inline float wgb_rcp_approx(float b) { return 1.0f / b; }
