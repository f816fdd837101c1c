// Device check of wgb_div_shared (wgb_raster.cuh): the hoisted-reciprocal division must equal __fdiv_rn bit for bit
// wherever its guard admits the operands.   nvcc -gencode arch=compute_100a,code=sm_100a -o div_check div_check.cu && ./div_check
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef float f32; typedef unsigned int u32; typedef unsigned long long u64;
#define WGB_DEV static __device__ __forceinline__
#define WGB_DIV_LO 8.6736174e-19f
#define WGB_DIV_HI 1.1529215e+18f
WGB_DEV f32 wgb_div_rcp(f32 b) {
    if (!(fabsf(b) >= WGB_DIV_LO && fabsf(b) <= WGB_DIV_HI)) return 0.0f;
    f32 r0;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(b));
    return __fmaf_rn(r0, __fmaf_rn(-b, r0, 1.0f), r0);
}
WGB_DEV f32 wgb_div_shared(f32 a, f32 b, f32 r) {
    const f32 q = __fmaf_rn(a, r, 0.0f);
    return __fmaf_rn(r, __fmaf_rn(-b, q, a), q);
}
WGB_DEV u64 mix(u64 x) { x += 0x9E3779B97F4A7C15ull; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull; x = (x ^ (x >> 27)) * 0x94D049BB133111EBull; return x ^ (x >> 31); }
// mode 0: random bit patterns inside the window; 1: a = k * b +- ulps (near-exact quotients); 2: a near a rounding
// boundary of the quotient (q + half ulp) * b; 3: small integers / shoelace-like magnitudes
__global__ void check(u64 seed, int mode, u64 per_thread, unsigned long long* bad, unsigned long long* tested, float* first) {
    u64 s = mix(seed ^ ((u64)blockIdx.x * blockDim.x + threadIdx.x) * 0x100000001B3ull);
    u64 nbad = 0, ntest = 0;
    for (u64 it = 0; it < per_thread; it++) {
        s = mix(s);
        f32 a, b;
        if (mode == 0) {
            u32 ea = 67u + (u32)((s >> 8) % 121u), eb = 67u + (u32)((s >> 16) % 121u);
            a = __uint_as_float(((u32)(s >> 63) << 31) | (ea << 23) | (u32)(s & 0x7FFFFFu));
            u64 t = mix(s ^ 0xABCDEFull);
            b = __uint_as_float(((u32)(t >> 63) << 31) | (eb << 23) | (u32)(t & 0x7FFFFFu));
        } else if (mode == 1) {
            b = __uint_as_float((127u << 23) | (u32)(s & 0x7FFFFFu)) * (f32)(1 + ((s >> 24) & 1023));
            const f32 k = (f32)(1 + ((s >> 34) & 0xFFFF));
            a = __fmul_rn(k, b);
            a = __uint_as_float(__float_as_uint(a) + (int)((s >> 50) & 7) - 3);
            if ((s >> 60) & 1) a = -a;
        } else if (mode == 2) {
            b = __uint_as_float((127u << 23) | (u32)(s & 0x7FFFFFu));
            const f32 q = __uint_as_float(((100u + (u32)((s >> 23) & 63u)) << 23) | (u32)((s >> 29) & 0x7FFFFFu));
            const double mid = ((double)q + (double)__uint_as_float(__float_as_uint(q) + 1)) * 0.5 * (double)b;
            a = (f32)mid;
            a = __uint_as_float(__float_as_uint(a) + (int)((s >> 56) & 3) - 1);
        } else {
            a = (f32)((int)((s >> 8) & 0xFFFFFF) - 0x800000) * (((s >> 40) & 1) ? 0.03125f : 1.0f);
            b = (f32)((int)((s >> 32) & 0x3FFFFF) - 0x200000) * (((s >> 41) & 1) ? 0.0078125f : 1.0f);
        }
        const f32 r = wgb_div_rcp(b);
        if (r == 0.0f || !(fabsf(a) >= WGB_DIV_LO && fabsf(a) <= WGB_DIV_HI)) continue;
        ntest++;
        const f32 want = __fdiv_rn(a, b), got = wgb_div_shared(a, b, r);
        if (__float_as_uint(want) != __float_as_uint(got)) { if (atomicAdd(bad, 1ull) == 0ull) { first[0] = a; first[1] = b; first[2] = want; first[3] = got; } nbad++; }
    }
    atomicAdd(tested, ntest);
}
int main() {
    unsigned long long *bad, *tested; float* first;
    cudaMallocManaged(&bad, 8); cudaMallocManaged(&tested, 8); cudaMallocManaged(&first, 16);
    int rc = 0;
    for (int mode = 0; mode < 4; mode++) {
        *bad = 0; *tested = 0;
        check<<<148 * 8, 256>>>(0x1234567ull + mode, mode, 1ull << 15, bad, tested, first);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("launch failed\n"); return 2; }
        printf("mode %d: tested %llu mismatches %llu", mode, *tested, *bad);
        if (*bad) { printf("  first: a=%a b=%a want=%a got=%a", first[0], first[1], first[2], first[3]); rc = 1; }
        printf("\n");
    }
    return rc;
}
