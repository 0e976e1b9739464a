// BEV tokeniser, tcgen05 variant (see tokens.cu for the operator and DESIGN.md 4b for the numbers).
//
// The FFMA2 and mma.sync projections of tokens.cu are limited by operand delivery: every warp drags the projection matrix
// through the LSU for a handful of cells.  Here the matrix stays in shared memory behind UMMA descriptors and the active
// cells of the WHOLE batch are compacted into 128-row tiles, so the projection is a plain [cells x C] x [C x d] GEMM on the
// 5th-generation tensor cores with the accumulator in tensor memory:
//
//   k_tok_stream_list   one pass over the index map: cells whose 3x3 window is empty get `background + PE` streamed to
//                       every frame (the HBM-bound part, ~82 % of the output on a pillar canvas); the other (frame, cell)
//                       pairs are appended to a global list.
//   k_tok_umma          persistent, one CTA per SM.  W (tf32 hi and lo halves, K-major, 128-byte swizzle: the image is
//                       prepared once by k_tok_wimg) is loaded into shared memory once.  Per 128-pair tile: refine
//                       (depthwise 3x3 + GELU from the pillar rows) writes A_hi / A_lo straight into the swizzled layout,
//                       one thread issues 3 x C/8 tcgen05.mma.kind::tf32 (M = 128, N = d, K = 8:
//                       a_lo.w_hi + a_hi.w_lo + a_hi.w_hi, fp32-accurate), tcgen05.commit -> mbarrier; the epilogue reads
//                       the accumulator with tcgen05.ld (one thread = one cell = one TMEM lane), so LayerNorm is a
//                       per-thread loop plus one shared-memory exchange between the two column halves; + PE; store.
//
// Shapes: C in {32, 64}, d in {128, 256} (W_hi + W_lo + A_hi + A_lo <= 192 KB of shared memory, N <= 256 per MMA).
#include "common.cuh"

namespace pillars {

namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kUT = 256;        // threads of the streaming kernel
constexpr int kUmmaR = 768;     // refine group of the tcgen05 kernel: latency-bound row gathers, so it gets most of the warps
                                // (measured on 16 x 512^2, d = 256: R/E = 512/512 1473 us, 768/256 1365 us, 896/128 1372 us)
constexpr int kUmmaE = 256;     // epilogue group: 4 TMEM lane quarters x kParts column parts
constexpr int kUmmaT = kUmmaR + kUmmaE;
constexpr int kParts = kUmmaE / 128;  // column parts of the epilogue group
constexpr int kUW = kUT / 32;
constexpr int kTileM = 128;     // pairs per GEMM tile = TMEM lanes
constexpr int kTileX = 32;
constexpr int kCellsPerWarp = kTileX / kUW;
constexpr int kFrameChunk = 16;
constexpr uint32_t kNoEntry = 0xFFFFFFFFu;

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }

struct StreamListParams {
    const int32_t *cell_row;  // [nb, h, w]
    int nb, h, w, d;
    const float *pe, *bg;
    float *out;
    uint32_t *list;           // [nb * h * w] worst case
    uint32_t *count;
};

// -------------------------------------------------------------------------------------------------------------------
// pass over the index map: stream the input-independent tokens, list the rest.  NQ = d / 128.
// -------------------------------------------------------------------------------------------------------------------
template <int NQ>
__global__ void __launch_bounds__(kUT) k_tok_stream_list(const __grid_constant__ StreamListParams p)
{
    __shared__ int32_t s_map[kFrameChunk][3][kTileX + 2];
    __shared__ uint32_t s_act[kFrameChunk];
    __shared__ uint32_t s_list[kTileX * kFrameChunk];
    __shared__ uint32_t s_count, s_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int d = p.d, w = p.w, h = p.h;
    const int x0 = blockIdx.x * kTileX, y = blockIdx.y;
    const size_t plane = static_cast<size_t>(h) * w;

    for (int b0 = 0; b0 < p.nb; b0 += kFrameChunk) {
        const int nbb = min(kFrameChunk, p.nb - b0);
        __syncthreads();
        if (tid == 0) s_count = 0;
        for (int i = tid; i < nbb * 3 * (kTileX + 2); i += kUT) {
            const int bb = i / (3 * (kTileX + 2)), r = i - bb * 3 * (kTileX + 2);
            const int dy = r / (kTileX + 2), dx = r - dy * (kTileX + 2);
            const int yy = y + dy - 1, xx = x0 + dx - 1;
            int32_t v = -1;
            if (yy >= 0 && yy < h && xx >= 0 && xx < w) v = __ldg(p.cell_row + (static_cast<size_t>(b0 + bb) * h + yy) * w + xx);
            s_map[bb][dy][dx] = v;
        }
        __syncthreads();
        for (int i = tid; i < nbb * kTileX; i += kUT) {
            const int bb = i >> 5, t = i & 31;
            bool any = false;
#pragma unroll
            for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) any |= s_map[bb][dy][t + dx] >= 0;
            const unsigned m = __ballot_sync(kFull, any);
            if (t == 0) s_act[bb] = m;
        }
        __syncthreads();
        for (int tt = 0; tt < kCellsPerWarp; ++tt) {
            const int t = warp * kCellsPerWarp + tt, x = x0 + t;
            if (x >= w) break;
            const size_t cell = static_cast<size_t>(y) * w + x;
            const unsigned amask = __ballot_sync(kFull, lane < nbb && ((s_act[lane & (kFrameChunk - 1)] >> t) & 1u));
            if (amask) {
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(&s_count, static_cast<uint32_t>(__popc(amask)));
                base = __shfl_sync(kFull, base, 0);
                if ((amask >> lane) & 1u)
                    s_list[base + __popc(amask & ((1u << lane) - 1u))] = static_cast<uint32_t>((b0 + lane) * plane + cell);
            }
            if (__popc(amask) == nbb) continue;
            float4 v[NQ];
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                const float4 a = __ldg(reinterpret_cast<const float4 *>(p.bg) + lane + 32 * q);
                const float4 e = __ldg(reinterpret_cast<const float4 *>(p.pe + cell * d) + lane + 32 * q);
                v[q] = make_float4(a.x + e.x, a.y + e.y, a.z + e.z, a.w + e.w);
            }
            float4 *dst = reinterpret_cast<float4 *>(p.out + (static_cast<size_t>(b0) * plane + cell) * d) + lane;
            const size_t frame_step = plane * (d >> 2);
            for (int bb = 0; bb < nbb; ++bb, dst += frame_step) {
                if ((amask >> bb) & 1u) continue;
#pragma unroll
                for (int q = 0; q < NQ; ++q) __stcs(dst + 32 * q, v[q]);
            }
        }
        __syncthreads();
        const uint32_t n = s_count;
        if (n) {
            if (tid == 0) s_base = atomicAdd(p.count, n);
            __syncthreads();
            for (uint32_t i = tid; i < n; i += kUT) p.list[s_base + i] = s_list[i];
        }
    }
}

// -------------------------------------------------------------------------------------------------------------------
// tcgen05 plumbing (PTX as in the CUTLASS / DeepGEMM sm_100 headers: cute/arch/mma_sm100_desc.hpp, mma_sm100_umma.hpp)
// -------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// shared-memory matrix descriptor: K-major tile, 128-byte swizzle, 8-row groups 1024 B apart (SBO), LBO unused, version 1
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr)
{
    return static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4) | (static_cast<uint64_t>(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor for kind::tf32: D fp32, A/B tf32, both K-major, M = 128
__device__ __forceinline__ uint32_t umma_idesc_tf32(int n) { return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(n >> 3) << 17) | ((128u >> 4) << 24); }

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0), "r"(0), "r"(0), "r"(0)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t mbar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory"); }
// bounded wait: try_wait parks the warp in hardware (no issue slots burnt while the other group works) and returns after a
// system-defined time limit at the latest; a bounded number of retries turns a set-up mistake into an error code, not a hang
__device__ __forceinline__ bool mbar_wait(uint32_t mbar, uint32_t parity)
{
    for (uint32_t spin = 0; spin < (1u << 16); ++spin) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(ok)
                     : "r"(mbar), "r"(parity)
                     : "memory");
        if (ok) return true;
    }
    return false;
}

// 32 consecutive accumulator columns of the thread's TMEM lane
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32])
{
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, "
        "%18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\ttcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
}

struct UmmaParams {
    const float *feats;
    const int32_t *cell_row;
    int nb, h, w, c, d;
    const float *dw_w, *dw_b, *pb, *gamma, *beta;
    float eps;
    const float *pe;
    const uint4 *wimg;        // [2 (hi, lo)][c / 32][d][128 B], swizzled: the exact shared-memory image
    const uint32_t *list, *count;
    float *out;
};

// tf32 split of a float4: hi keeps sign, exponent and 10 mantissa bits (what the tensor core reads), lo = x - hi (exact)
__device__ __forceinline__ void split4(const float4 v, float4 &hi, float4 &lo)
{
    hi.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); lo.x = v.x - hi.x;
    hi.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u); lo.y = v.y - hi.y;
    hi.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); lo.z = v.z - hi.z;
    hi.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u); lo.w = v.w - hi.w;
}

__device__ __forceinline__ void bar_named(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t mbar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mbar) : "memory"); }

// Warp-specialised, two tiles in flight per CTA (1024 threads at 64 registers):
//   warps 0-23  (group R, 768 threads)  metadata + refine of tile i+1 into the A buffers; thread 0 issues the MMAs
//   warps 24-31 (group E, 256 threads)  LayerNorm + PE + stores of tile i out of tensor memory
// The accumulator is double buffered in TMEM (2 x d columns), so MMA(i+1) may run while E still reads tile i; the A buffers
// are single: refine(i+1) starts when the commit of MMA(i) has arrived (a_free).  mbarriers: a_free (1 arrival: commit),
// acc_ready[2] (1: commit), acc_free[2] (kUmmaE: every E thread after its last TMEM read of the tile).
template <int C, int D>
__global__ void __launch_bounds__(kUmmaT, 1) k_tok_umma(const __grid_constant__ UmmaParams p)
{
    // channel counts are compile-time: shared-memory addressing of the weights and of the swizzled tiles folds into immediates
    extern __shared__ __align__(1024) uint8_t s_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int c = C, d = D;
    const int w = p.w, h = p.h;
    constexpr int kb_n = C >> 5, quads = C >> 2, qshift = quads == 16 ? 4 : 3;
    const uint32_t w_half = static_cast<uint32_t>(kb_n) * d * 128u;       // bytes of W_hi (= W_lo)
    const uint32_t a_half = static_cast<uint32_t>(kb_n) * kTileM * 128u;  // bytes of A_hi (= A_lo)
    uint8_t *const s_al = s_raw + ((1024u - (smem_u32(s_raw) & 1023u)) & 1023u);  // swizzle atoms need 1024-byte alignment
    uint8_t *const s_w = s_al;                        // W_hi | W_lo
    uint8_t *const s_a = s_al + 2 * w_half;           // A_hi | A_lo
    float *const s_dw = reinterpret_cast<float *>(s_a + 2 * a_half);  // [10][c]
    float *const s_vec = s_dw + 10 * c;                                // pb | gamma | beta, [3][d]
    int32_t *const s_nb_all = reinterpret_cast<int32_t *>(s_vec + 3 * d);                 // [2][128][9]
    uint32_t *const s_ent_all = reinterpret_cast<uint32_t *>(s_nb_all + 2 * kTileM * 9);  // [2][128]
    float *const s_stat = reinterpret_cast<float *>(s_ent_all + 2 * kTileM);              // [2][128] (E group, column halves)
    uint64_t *const s_mbar = reinterpret_cast<uint64_t *>(s_stat + 2 * kParts * kTileM);      // a_free, acc_ready[2], acc_free[2]
    uint32_t *const s_tmem = reinterpret_cast<uint32_t *>(s_mbar + 5);
    volatile uint32_t *const s_abort = s_tmem + 1;
    const uint32_t mb_a_free = smem_u32(s_mbar), mb_ready0 = smem_u32(s_mbar + 1), mb_free0 = smem_u32(s_mbar + 3);
    const uint32_t tmem_cols = 2u * static_cast<uint32_t>(d);

    // ---- one-time setup ------------------------------------------------------------------------------------------------------
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 32) {
        mbar_init(mb_a_free, 1);
        mbar_init(mb_ready0, 1);
        mbar_init(mb_ready0 + 8, 1);
        mbar_init(mb_free0, kUmmaE);
        mbar_init(mb_free0 + 8, kUmmaE);
        *s_abort = 0u;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {
        const uint4 *src = p.wimg;
        uint4 *dst = reinterpret_cast<uint4 *>(s_w);
        for (uint32_t i = tid; i < (2 * w_half) >> 4; i += kUmmaT) dst[i] = __ldg(src + i);
    }
    for (int i = tid; i < 9 * c; i += kUmmaT) {
        const int ch = i / 9, k = i - ch * 9;
        s_dw[k * c + ch] = __ldg(p.dw_w + i);
    }
    for (int i = tid; i < c; i += kUmmaT) s_dw[9 * c + i] = __ldg(p.dw_b + i);
    for (int i = tid; i < d; i += kUmmaT) {
        s_vec[i] = __ldg(p.pb + i);
        s_vec[d + i] = __ldg(p.gamma + i);
        s_vec[2 * d + i] = __ldg(p.beta + i);
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;
    uint32_t n_total = __ldg(p.count);
    const uint32_t idesc = umma_idesc_tf32(d);
    const size_t plane = static_cast<size_t>(h) * w;
    const uint32_t n_tiles = (n_total + kTileM - 1) / kTileM;
    const auto report = [&](uint32_t code) {  // bounded waits never expire in a correct run; leave a trace if one does
        *s_abort = 1u;
        atomicExch(const_cast<uint32_t *>(p.count) + 1, 0xDEAD0000u | code);
    };

    if (warp < kUmmaR / 32) {
        // ================================ group R: metadata, refine, MMA issue ===================================================
        const int rt = tid;  // 0..255
        auto load_meta = [&](uint32_t tile, int buf) {
            if (rt < kTileM) {
                const uint32_t idx = tile * kTileM + rt;
                const uint32_t e = (tile < n_tiles && idx < n_total) ? __ldg(p.list + idx) : kNoEntry;
                s_ent_all[buf * kTileM + rt] = e;
                if (e != kNoEntry) {
                    const uint32_t b = e / static_cast<uint32_t>(plane), cell = e - b * static_cast<uint32_t>(plane);
                    const int yy0 = static_cast<int>(cell / w), xx0 = static_cast<int>(cell - static_cast<uint32_t>(yy0) * w);
                    int32_t *nb = s_nb_all + (buf * kTileM + rt) * 9;
#pragma unroll
                    for (int k = 0; k < 9; ++k) {
                        const int yy = yy0 + k / 3 - 1, xx = xx0 + k % 3 - 1;
                        int32_t r = -1;
                        if (yy >= 0 && yy < h && xx >= 0 && xx < w) r = __ldg(p.cell_row + (static_cast<size_t>(b) * h + yy) * w + xx);
                        nb[k] = r >= 0 ? r * quads : -1;  // float4 index of the row start: the refine adds the quad
                        if (r >= 0) {
                            const char *row = reinterpret_cast<const char *>(p.feats + static_cast<size_t>(r) * c);
                            for (int o = 0; o < c * 4; o += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(row + o));
                        }
                    }
                }
            }
        };
        load_meta(blockIdx.x, 0);
        bar_named(1, kUmmaR);
        uint32_t it = 0;
        for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int cur = it & 1;
            const int32_t *const s_nb = s_nb_all + cur * kTileM * 9;
            const uint32_t *const s_ent = s_ent_all + cur * kTileM;
            // the A buffers are free once the previous tile's MMAs have completed
            if (it > 0 && !mbar_wait(mb_a_free, (it - 1) & 1u)) { report(1); break; }
            for (int item = rt; item < kTileM * quads; item += kUmmaR) {
                const int row = item >> qshift, qd = item & (quads - 1);  // quads is 8 or 16
                float4 act = make_float4(0.f, 0.f, 0.f, 0.f);
                if (s_ent[row] != kNoEntry) {
                    // all nine row gathers are requested before the first one is consumed
                    float4 f[9];
#pragma unroll
                    for (int k = 0; k < 9; ++k) {
                        const int32_t r = s_nb[row * 9 + k];
                        f[k] = r >= 0 ? __ldg(reinterpret_cast<const float4 *>(p.feats) + (r + qd)) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                    float4 acc = *reinterpret_cast<const float4 *>(s_dw + 9 * c + 4 * qd);
#pragma unroll
                    for (int k = 0; k < 9; ++k) {
                        const float4 wk = *reinterpret_cast<const float4 *>(s_dw + k * c + 4 * qd);
                        acc.x = fmaf(wk.x, f[k].x, acc.x);
                        acc.y = fmaf(wk.y, f[k].y, acc.y);
                        acc.z = fmaf(wk.z, f[k].z, acc.z);
                        acc.w = fmaf(wk.w, f[k].w, acc.w);
                    }
                    act = make_float4(gelu_erf(acc.x), gelu_erf(acc.y), gelu_erf(acc.z), gelu_erf(acc.w));
                }
                float4 hi, lo;
                split4(act, hi, lo);
                const uint32_t off = static_cast<uint32_t>(qd >> 3) * (kTileM * 128u) + static_cast<uint32_t>(row) * 128u +
                                     ((static_cast<uint32_t>(qd & 7) ^ static_cast<uint32_t>(row & 7)) << 4);
                *reinterpret_cast<float4 *>(s_a + off) = hi;
                *reinterpret_cast<float4 *>(s_a + a_half + off) = lo;
            }
            fence_async_smem();  // generic-proxy stores -> visible to the tensor core's async proxy
            tc_fence_before();
            bar_named(1, kUmmaR);
            if (rt == 0) {
                // the accumulator stage must have been drained by the epilogue of tile it - 2
                bool ok = true;
                if (it >= 2) ok = mbar_wait(mb_free0 + 8u * cur, ((it >> 1) - 1) & 1u);
                if (!ok) report(2);
                tc_fence_after();
                const uint32_t a_hi = smem_u32(s_a), a_lo = a_hi + a_half, w_hi = smem_u32(s_w), w_lo = w_hi + w_half;
                const uint32_t tacc = tmem_base + static_cast<uint32_t>(cur) * static_cast<uint32_t>(d);
                uint32_t accumulate = 0;
                for (int kb = 0; kb < kb_n; ++kb) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint32_t ao = static_cast<uint32_t>(kb) * (kTileM * 128u) + k * 32u;
                        const uint32_t bo = static_cast<uint32_t>(kb) * (static_cast<uint32_t>(d) * 128u) + k * 32u;
                        umma_tf32(tacc, umma_desc_sw128(a_lo + ao), umma_desc_sw128(w_hi + bo), idesc, accumulate);
                        umma_tf32(tacc, umma_desc_sw128(a_hi + ao), umma_desc_sw128(w_lo + bo), idesc, 1u);
                        umma_tf32(tacc, umma_desc_sw128(a_hi + ao), umma_desc_sw128(w_hi + bo), idesc, 1u);
                        accumulate = 1u;
                    }
                }
                umma_commit(mb_a_free);
                umma_commit(mb_ready0 + 8u * cur);
            }
            load_meta(tile + gridDim.x, cur ^ 1);  // next tile's lookups run under this tile's MMAs
            bar_named(1, kUmmaR);
        }
    } else {
        // ================================ group E: LayerNorm + PE + stores out of tensor memory ===================================
        const int et = tid - kUmmaR;              // 0..255 (kUmmaR is a multiple of 128: the lane quarter is still warp % 4)
        const int ew = et >> 5;                   // 0..7; TMEM lane quarter = warp % 4 (8 % 4 == 0, so ew % 4 too)
        const int row = 32 * (ew & 3) + lane, hf = ew >> 2;
        const int half_cols = d / kParts, col0 = hf * half_cols;
        uint32_t it = 0;
        for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int cur = it & 1;
            const uint32_t idx = tile * kTileM + row;
            const uint32_t e = idx < n_total ? __ldg(p.list + idx) : kNoEntry;
            const bool live = e != kNoEntry;  // tcgen05.ld is warp-collective: every lane loads, only live rows store
            const uint32_t b = live ? e / static_cast<uint32_t>(plane) : 0u, cell = live ? e - b * static_cast<uint32_t>(plane) : 0u;
            const float4 *pe4 = reinterpret_cast<const float4 *>(p.pe + static_cast<size_t>(cell) * d + col0);
            float4 *dst = reinterpret_cast<float4 *>(p.out + static_cast<size_t>(live ? e : 0u) * d + col0);
            if (live)
                for (int o = 0; o < half_cols * 4; o += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char *>(pe4) + o));
            if (!mbar_wait(mb_ready0 + 8u * cur, (it >> 1) & 1u)) { report(3); break; }
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(32 * (ew & 3)) << 16) + static_cast<uint32_t>(cur * d + col0);
            float v[32];
            // LayerNorm statistics in ONE pass over the accumulator: sums of (x - c) and (x - c)^2 around the pivot c = the
            // thread's first value (the shifted-data form: no cancellation as long as c lies inside the data), then the
            // column parts are merged with the pairwise mean / M2 update.
            float s1 = 0.f, s2 = 0.f, piv = 0.f;
#pragma unroll 1
            for (int ch = 0; ch < half_cols; ch += 32) {
                tmem_ld32(taddr + ch, v);
                if (ch == 0) piv = v[0] + s_vec[col0];
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 pb = *reinterpret_cast<const float4 *>(s_vec + col0 + ch + j);
                    const float x0 = v[j] + pb.x - piv, x1 = v[j + 1] + pb.y - piv, x2 = v[j + 2] + pb.z - piv, x3 = v[j + 3] + pb.w - piv;
                    s1 += (x0 + x1) + (x2 + x3);
                    s2 = fmaf(x0, x0, s2); s2 = fmaf(x1, x1, s2); s2 = fmaf(x2, x2, s2); s2 = fmaf(x3, x3, s2);
                }
            }
            const float n_part = static_cast<float>(half_cols);
            float mean = piv + s1 / n_part;          // mean of this thread's columns
            float m2 = s2 - s1 * s1 / n_part;        // sum of squared deviations from it
            if constexpr (kParts > 1) {
                s_stat[hf * kTileM + row] = mean;
                s_stat[(kParts + hf) * kTileM + row] = m2;
                bar_named(2, kUmmaE);
                float msum = 0.f;
#pragma unroll
                for (int pp = 0; pp < kParts; ++pp) msum += s_stat[pp * kTileM + row];
                const float mall = msum / static_cast<float>(kParts);
                float m2all = 0.f;
#pragma unroll
                for (int pp = 0; pp < kParts; ++pp) {
                    const float dm = s_stat[pp * kTileM + row] - mall;
                    m2all += s_stat[(kParts + pp) * kTileM + row] + n_part * dm * dm;
                }
                bar_named(2, kUmmaE);
                mean = mall;
                m2 = m2all;
            }
            const float rstd = 1.f / sqrtf(fmaxf(m2, 0.f) / static_cast<float>(d) + p.eps);
#pragma unroll 1
            for (int ch = 0; ch < half_cols; ch += 32) {
                tmem_ld32(taddr + ch, v);
                if (ch + 32 >= half_cols) {  // last read of this accumulator stage: hand it back to the MMA issuer
                    tc_fence_before();
                    mbar_arrive(mb_free0 + 8u * cur);
                }
                if (live) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const int col = col0 + ch + j;
                        const float4 pb = *reinterpret_cast<const float4 *>(s_vec + col);
                        const float4 ga = *reinterpret_cast<const float4 *>(s_vec + d + col);
                        const float4 be = *reinterpret_cast<const float4 *>(s_vec + 2 * d + col);
                        const float4 pe = __ldg(pe4 + ((ch + j) >> 2));
                        float4 o;
                        o.x = fmaf((v[j] + pb.x - mean) * rstd, ga.x, be.x) + pe.x;
                        o.y = fmaf((v[j + 1] + pb.y - mean) * rstd, ga.y, be.y) + pe.y;
                        o.z = fmaf((v[j + 2] + pb.z - mean) * rstd, ga.z, be.z) + pe.z;
                        o.w = fmaf((v[j + 3] + pb.w - mean) * rstd, ga.w, be.w) + pe.w;
                        __stcs(dst + ((ch + j) >> 2), o);
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
}

// W image: element (n, k) of Wp = wt[k][n], half hf (0 hi, 1 lo), at [(hf * kb_n + k / 32) * d + n] * 128 B + 16-byte chunk
// ((k % 32) / 4) ^ (n % 8), position k % 4 -- the K-major 128-byte-swizzled tile the descriptor above describes.
__global__ void k_tok_wimg(const float *__restrict__ wt, int c, int d, float *__restrict__ img)
{
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;  // one thread per (k quad, n)
    const int quads = c >> 2, kb_n = c >> 5;
    if (i >= static_cast<int64_t>(quads) * d) return;
    const int n = static_cast<int>(i % d), qd = static_cast<int>(i / d);
    const size_t half = static_cast<size_t>(kb_n) * d * 32;  // floats
    const size_t off = (static_cast<size_t>(qd >> 3) * d + n) * 32 + ((static_cast<size_t>(qd & 7) ^ static_cast<size_t>(n & 7)) << 2);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float v = wt[static_cast<size_t>(4 * qd + j) * d + n];
        const float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
        img[off + j] = hi;
        img[half + off + j] = v - hi;
    }
}

size_t umma_smem_bytes(int c, int d)
{
    const size_t kb_n = static_cast<size_t>(c) >> 5;
    return 2 * kb_n * d * 128 + 2 * kb_n * kTileM * 128 + sizeof(float) * (10 * static_cast<size_t>(c) + 3 * static_cast<size_t>(d)) +
           2 * (sizeof(int32_t) * kTileM * 9 + sizeof(uint32_t) * kTileM) + sizeof(float) * 2 * kParts * kTileM + 5 * 8 + 16;
}

}  // namespace

bool tokens_umma_supported(int c, int d) { return (c == 32 || c == 64) && (d == 128 || d == 256) && umma_smem_bytes(c, d) <= 227 * 1024; }

cudaError_t launch_tokens_wimg(const float *wt, int c, int d, float *img, cudaStream_t st)
{
    const int64_t n = static_cast<int64_t>(c >> 2) * d;
    k_tok_wimg<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(wt, c, d, img);
    note_launch();
    return cudaGetLastError();
}

cudaError_t launch_bev_tokens_umma(const TokenizerDev &tk, const float *feats, const int32_t *cell_row, int nb, int h, int w,
                                   float *out, uint32_t *list, uint32_t *count, cudaStream_t st)
{
    if (nb == 0 || h == 0 || w == 0) return cudaSuccess;
    if (static_cast<uint64_t>(nb) * h * w >= 0xFFFFFFFFull) return cudaErrorInvalidValue;  // pairs are 32-bit
    cudaError_t e = cudaMemsetAsync(count, 0, sizeof(uint32_t), st);
    note_launch();
    if (e != cudaSuccess) return e;
    StreamListParams sp{};
    sp.cell_row = cell_row; sp.nb = nb; sp.h = h; sp.w = w; sp.d = tk.d; sp.pe = tk.pe; sp.bg = tk.bg; sp.out = out;
    sp.list = list; sp.count = count;
    const dim3 grid(static_cast<unsigned>((w + kTileX - 1) / kTileX), static_cast<unsigned>(h));
    // Measured: running the streaming pass beside the tcgen05 kernel (helper stream, shared-memory-free streaming CTAs sized to
    // what one tcgen05 CTA leaves of an SM) is slower (1.59 ms) than one after the other (1.42 ms): the streaming pass needs
    // the whole SM's warps to saturate HBM.
    if (tk.d == 128) k_tok_stream_list<1><<<grid, kUT, 0, st>>>(sp);
    else k_tok_stream_list<2><<<grid, kUT, 0, st>>>(sp);
    note_launch();
    if ((e = cudaGetLastError()) != cudaSuccess) return e;

    UmmaParams up{};
    up.feats = feats; up.cell_row = cell_row; up.nb = nb; up.h = h; up.w = w; up.c = tk.c; up.d = tk.d;
    up.dw_w = tk.dw_w; up.dw_b = tk.dw_b; up.pb = tk.pb; up.gamma = tk.gamma; up.beta = tk.beta; up.eps = tk.eps; up.pe = tk.pe;
    up.wimg = reinterpret_cast<const uint4 *>(tk.wimg); up.list = list; up.count = count; up.out = out;
    const size_t smem = umma_smem_bytes(tk.c, tk.d) + 1024;  // slack for the 1024-byte alignment of the tiles
    // (function attributes are per device and a process may drive more than one: set every time, it is cheap)
    const int sms = current_sm_count();
    {
        cudaError_t e;
        if ((e = cudaFuncSetAttribute(k_tok_umma<64, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)) != cudaSuccess) return e;
        if ((e = cudaFuncSetAttribute(k_tok_umma<64, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)) != cudaSuccess) return e;
        if ((e = cudaFuncSetAttribute(k_tok_umma<32, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)) != cudaSuccess) return e;
        if ((e = cudaFuncSetAttribute(k_tok_umma<32, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)) != cudaSuccess) return e;
    }
    if (tk.c == 64 && tk.d == 256) k_tok_umma<64, 256><<<sms, kUmmaT, smem, st>>>(up);
    else if (tk.c == 64) k_tok_umma<64, 128><<<sms, kUmmaT, smem, st>>>(up);
    else if (tk.d == 256) k_tok_umma<32, 256><<<sms, kUmmaT, smem, st>>>(up);
    else k_tok_umma<32, 128><<<sms, kUmmaT, smem, st>>>(up);
    note_launch();
    return cudaGetLastError();
}

}  // namespace pillars
