// Checks wgb_span_div (wgb_raster.cuh) against 32-bit truncating division over the whole operand range the span
// DDA can produce: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o span_div_check span_div_check.cu && ./span_div_check
#include <cstdio>
__device__ int wgb_span_div(int dx, int i, int dy) {
    const int num = dx * i;
    const int a = abs(num);
    int q = __float2int_rz(__fdividef(__int2float_rn(a), __int2float_rn(dy)));
    const int r = a - q * dy;
    if (r < 0) q -= 1; else if (r >= dy) q += 1;
    return num < 0 ? -q : q;
}
__global__ void k(unsigned long long* bad) {
    // exhaustive-ish: dy in [1, 32767] strided, dx in [-32767, 32767] strided, all i in [0, dy] strided
    const int dy = blockIdx.x * 7 + 1;
    for (int dx = -32767 + (int)threadIdx.x; dx <= 32767; dx += (int)blockDim.x * 3)
        for (int i = 0; i <= dy; i += (dy > 64 ? dy / 61 : 1)) {
            if (wgb_span_div(dx, i, dy) != (dx * i) / dy) atomicAdd(bad, 1ull);
            const int j = dy - (i % 5);          // near the end too
            if (j >= 0 && wgb_span_div(dx, j, dy) != (dx * j) / dy) atomicAdd(bad, 1ull);
        }
}
int main() { unsigned long long* bad; cudaMallocManaged(&bad, 8); *bad = 0; k<<<4681, 256>>>(bad); cudaDeviceSynchronize(); printf("mismatches: %llu (%s)\n", *bad, cudaGetErrorString(cudaGetLastError())); return 0; }
