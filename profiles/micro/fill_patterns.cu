// Micro-benchmark: how fast can a kernel WRITE a [16][64][512*512] fp32 canvas (1 GiB) on B200, by access pattern?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fill_patterns fill_patterns.cu && ./fill_patterns
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ void st256(float *p) {
    asm volatile("st.global.v8.f32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(p), "f"(0.f) : "memory");
}
__device__ __forceinline__ void st256_na(float *p) {
    asm volatile("st.global.L1::no_allocate.v8.f32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(p), "f"(0.f) : "memory");
}
__device__ __forceinline__ void st128(float *p) {
    asm volatile("st.global.v4.f32 [%0], {%1,%1,%1,%1};" ::"l"(p), "f"(0.f) : "memory");
}
// A: linear, one 32 B store per thread, consecutive threads -> consecutive addresses, `per` stores per thread strided by grid
__global__ void k_linear256(float *p, size_t n8, int per) {
    size_t i = (size_t)blockIdx.x * blockDim.x * per + threadIdx.x;
    for (int k = 0; k < per; ++k, i += blockDim.x) if (i < n8) st256(p + i * 8);
}
__global__ void k_linear128(float *p, size_t n4, int per) {
    size_t i = (size_t)blockIdx.x * blockDim.x * per + threadIdx.x;
    for (int k = 0; k < per; ++k, i += blockDim.x) if (i < n4) st128(p + i * 4);
}
// C: canvas pattern: thread owns 8 cells, loops over 64 channels (stride plane)
__global__ void k_canvas256(float *p, int64_t plane, int tpp, int f, const int *idx) {
    const int b = blockIdx.x / tpp;
    const int64_t cell0 = (int64_t)(blockIdx.x % tpp) * (blockDim.x * 8) + threadIdx.x * 8;
    if (idx) { int4 v = __ldg((const int4 *)(idx + b * plane + cell0)); if (v.x == 12345) return; }
    float *d = p + ((int64_t)b * f) * plane + cell0;
#pragma unroll 8
    for (int c = 0; c < f; ++c) st256_na(d + c * plane);
}
// E: channel-group pattern: warp owns `cells` contiguous cells x 8 channels
__global__ void k_canvas_cg(float *p, int64_t plane, int f, int cells_per_cta) {
    // blockIdx.x -> (b, channel group of 8, cell tile)
    const int tiles = plane / cells_per_cta;
    const int t = blockIdx.x % tiles;
    const int cg = (blockIdx.x / tiles) % (f / 8);
    const int b = blockIdx.x / (tiles * (f / 8));
    float *d = p + ((int64_t)b * f + cg * 8) * plane + (int64_t)t * cells_per_cta;
    for (int c = 0; c < 8; ++c)
        for (int i = threadIdx.x * 8; i < cells_per_cta; i += blockDim.x * 8) st256_na(d + c * plane + i);
}
int main() {
    const int B = 16, F = 64; const int64_t plane = 512 * 512; const size_t n = (size_t)B * F * plane;
    float *p; int *idx; CK(cudaMalloc(&p, n * 4)); CK(cudaMalloc(&idx, B * plane * 4)); CK(cudaMemset(idx, 0xFF, B * plane * 4));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto timeit = [&](const char *name, auto launch) {
        float best = 1e9, sum = 0; for (int it = 0; it < 12; ++it) { cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (it >= 2) { sum += ms; if (ms < best) best = ms; } }
        printf("%-44s best %7.1f us  avg %7.1f us  %6.0f GB/s (best)\n", name, best * 1e3, sum / 10 * 1e3, n * 4 / (best * 1e-3) / 1e9);
        cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) printf("  ERR %s\n", cudaGetErrorString(e));
    };
    timeit("cudaMemsetAsync", [&] { cudaMemsetAsync(p, 0, n * 4); });
    for (int per : {1, 4, 16}) { char nm[64]; snprintf(nm, 64, "linear STG.256 per=%d bs=256", per); size_t n8 = n / 8; unsigned g = (n8 + 256 * per - 1) / (256 * per); timeit(nm, [&] { k_linear256<<<g, 256>>>(p, n8, per); }); }
    for (int per : {1, 4, 16}) { char nm[64]; snprintf(nm, 64, "linear STG.128 per=%d bs=256", per); size_t n4 = n / 4; unsigned g = (n4 + 256 * per - 1) / (256 * per); timeit(nm, [&] { k_linear128<<<g, 256>>>(p, n4, per); }); }
    for (int bs : {64, 256}) { int tpp = plane / (bs * 8); char nm[64]; snprintf(nm, 64, "canvas STG.256 bs=%d no index", bs); timeit(nm, [&] { k_canvas256<<<B * tpp, bs>>>(p, plane, tpp, F, nullptr); });
        snprintf(nm, 64, "canvas STG.256 bs=%d with index load", bs); timeit(nm, [&] { k_canvas256<<<B * tpp, bs>>>(p, plane, tpp, F, idx); }); }
    for (int cells : {2048, 8192, 32768}) { char nm[64]; snprintf(nm, 64, "chan-group 8ch x %d cells bs=256", cells); unsigned g = B * (F / 8) * (plane / cells); timeit(nm, [&] { k_canvas_cg<<<g, 256>>>(p, plane, F, cells); }); }
    return 0;
}
