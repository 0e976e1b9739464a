// Does a tcgen05 shared-memory descriptor (K-major, 128-byte swizzle) accept a start address that is NOT aligned to the
// 1024-byte swizzle atom, and a stride between 8-row groups that is not a multiple of 1024 B?  That is what an implicit-GEMM
// convolution needs to read all 3x3 taps out of ONE halo tile: tap (ky, kx) is the same buffer shifted by (ky * halo_w + kx)
// pixel rows of 128 B.  The buffer is written with the swizzle of the ABSOLUTE address (chunk ^= (addr >> 7) & 7).
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o umma_shifted_desc umma_shifted_desc.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t sbo_bytes, uint32_t base_off)
{
    return static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4) | (static_cast<uint64_t>(sbo_bytes >> 4) << 32) | (1ull << 46) |
           (static_cast<uint64_t>(base_off & 7u) << 49) | (2ull << 61);
}
__device__ __forceinline__ uint32_t idesc_tf32(int n) { return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(n >> 3) << 17) | ((128u >> 4) << 24); }

constexpr int kRows = 256;  // pixel rows of 128 B in the A buffer
constexpr int kN = 64;

// mode bit 0: base offset = (start >> 7) & 7 instead of 0
__global__ void __launch_bounds__(128) k_test(const float *a_rows /*[kRows][32]*/, const float *b_rows /*[kN][32]*/, int shift_rows,
                                              int sbo_rows, int mode, float *out /*[128][kN]*/)
{
    extern __shared__ __align__(1024) uint8_t raw[];
    uint8_t *s = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
    uint8_t *s_a = s, *s_b = s + kRows * 128;
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < kRows * 8; i += 128) {
        const int row = i >> 3, ch = i & 7;
        const uint32_t off = row * 128u + ((ch ^ (row & 7)) << 4);  // s_a is 1024-aligned: row & 7 == (addr >> 7) & 7
        *reinterpret_cast<float4 *>(s_a + off) = *reinterpret_cast<const float4 *>(a_rows + row * 32 + ch * 4);
    }
    for (int i = tid; i < kN * 8; i += 128) {
        const int row = i >> 3, ch = i & 7;
        *reinterpret_cast<float4 *>(s_b + row * 128u + ((ch ^ (row & 7)) << 4)) = *reinterpret_cast<const float4 *>(b_rows + row * 32 + ch * 4);
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(64u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(1u) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (tid == 0) {
        const uint32_t a0 = smem_u32(s_a) + shift_rows * 128u, b0 = smem_u32(s_b);
        const uint32_t bo = (mode & 1) ? ((a0 >> 7) & 7u) : 0u;
        for (int k = 0; k < 4; ++k) {
            const uint64_t ad = make_desc(a0 + k * 32u, sbo_rows * 128u, bo), bd = make_desc(b0 + k * 32u, 1024u, 0u);
            const uint32_t acc = k > 0;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}\n" ::"r"(tmem),
                         "l"(ad), "l"(bd), "r"(idesc_tf32(kN)), "r"(acc), "r"(0), "r"(0), "r"(0), "r"(0)
                         : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
    }
    {
        uint32_t ok = 0;
        for (int spin = 0; spin < (1 << 20) && !ok; ++spin)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                         : "=r"(ok) : "r"(smem_u32(&mbar)), "r"(0u) : "memory");
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < kN; c0 += 32) {
        uint32_t r[32];
        const uint32_t taddr = tmem + (static_cast<uint32_t>(32 * warp) << 16) + c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, "
            "%18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\ttcgen05.wait::ld.sync.aligned;"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
              "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
              "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
              "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr) : "memory");
        for (int j = 0; j < 32; ++j) out[tid * kN + c0 + j] = __uint_as_float(r[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u) : "memory");
}

int main()
{
    std::vector<float> a(kRows * 32), b(kN * 32);
    srand(1);
    // small integers: exact in tf32, so the comparison is exact
    for (auto &v : a) v = static_cast<float>(rand() % 17 - 8);
    for (auto &v : b) v = static_cast<float>(rand() % 13 - 6);
    float *da, *db, *dout;
    cudaMalloc(&da, a.size() * 4); cudaMalloc(&db, b.size() * 4); cudaMalloc(&dout, 128 * kN * 4);
    cudaMemcpy(da, a.data(), a.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(db, b.data(), b.size() * 4, cudaMemcpyHostToDevice);
    const size_t smem = kRows * 128 + kN * 128 + 1024;
    cudaFuncSetAttribute(k_test, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    std::vector<float> out(128 * kN);
    const int shifts[] = {0, 1, 2, 3, 7, 8, 9, 10, 11, 12, 21};
    const int sbos[] = {8, 10, 12, 16, 18};
    for (int mode = 0; mode < 2; ++mode)
        for (int sbo : sbos)
            for (int shift : shifts) {
                if (shift + 15 * sbo + 8 > kRows) continue;
                cudaMemset(dout, 0, out.size() * 4);
                k_test<<<1, 128, smem>>>(da, db, shift, sbo, mode, dout);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("mode %d sbo %d shift %d: CUDA error %s\n", mode, sbo, shift, cudaGetErrorString(e)); return 1; }
                cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
                int bad = 0;
                for (int m = 0; m < 128; ++m) {
                    const int row = shift + (m >> 3) * sbo + (m & 7);
                    for (int n = 0; n < kN; ++n) {
                        float ref = 0.f;
                        for (int k = 0; k < 32; ++k) ref += a[row * 32 + k] * b[n * 32 + k];
                        if (ref != out[m * kN + n]) ++bad;
                    }
                }
                printf("base_offset %s  group stride %2d rows  start shift %2d rows : %s (%d of %d wrong)\n",
                       mode ? "(addr>>7)&7" : "0          ", sbo, shift, bad ? "MISMATCH" : "ok", bad, 128 * kN);
            }
    return 0;
}
