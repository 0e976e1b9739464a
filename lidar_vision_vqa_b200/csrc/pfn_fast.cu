// Fast path of the per-pillar feature stage for the mainstream configuration (USE_ABSLOTE_XYZ, no WITH_DISTANCE,
// C <= 5 point channels, one PFN layer, F = 64).  Same function as pfn.cu / pillar_vfe.py:29-49,94-123, reorganised so that
// the FMA pipes do (almost) nothing but useful work:
//
//   * the points arrive as 32-byte records already grouped by pillar (k_place), so a CTA streams 256 consecutive list
//     positions instead of gathering;
//   * ONE THREAD PER POINT computes all 64 channels with the weights as constant-bank FFMA operands (they are a kernel
//     parameter), i.e. no weight registers, no shared-memory weight traffic, no lanes idling on short pillars;
//   * the linear layer is refactored around the pillar centre c:  with x' = x - c and m' = mean - c
//         W.[p, p_xyz - mean, p_xyz - c] = (W_p + W_cl + W_ce).x' + W_it.(i,t)  +  [ W_p.c - W_cl.m' ]
//     so a point costs 5 FMAs per channel and the bracket is one 6-FMA constant per (pillar, channel).  All per-point
//     terms are small numbers (|x'| <= voxel/2): no cancellation is introduced.  BatchNorm's scale is folded into W on the
//     host, and since "+ constant" and ReLU are monotone the max over the pillar's points is taken BEFORE them;
//   * the per-point results go through shared memory ([32 ch][289] per half) and 8 threads per pillar take the max.
//
// A CTA owns the pillars whose list STARTS inside its 256 positions; the tail of its last pillar (usually 0-2 points in
// the next chunk) is swept by warp 0.  Pillars over the cap (n > P) get their "first P by point index" threshold from a
// warp-wide radix select.
#include <cstdlib>

#include "common.cuh"

namespace pillars {

namespace {

constexpr int kTailW = 32; // tail positions per sweep (one warp)
constexpr int kHalf = 32;  // channels per pass
constexpr unsigned kFull = 0xffffffffu;

struct FastParams {
    const PointRecord *records;
    const Header *hdr;
    const uint32_t *pillar_key, *pillar_cnt, *frame_gstart, *frame_rowbase;
    int32_t *cell_row;
    float *pillar_features;
    int32_t *voxel_coords, *voxel_num_points;
    int64_t capacity;
    GridDev gd;
    float vsz[3], off[3];
    int idx_bits;
    FastWeights w;
};

__device__ __forceinline__ unsigned lanemask_le()
{
    unsigned m;
    asm("mov.u32 %0, %%lanemask_le;" : "=r"(m));
    return m;
}

template <int H, int kLd>
__device__ __forceinline__ void point_half(const FastParams &p, bool kept, float xp, float yp, float zp, float in, float tm,
                                           float *__restrict__ col)
{
    if (kept) {
#pragma unroll
        for (int c = 0; c < kHalf; ++c) {
            float a = xp * p.w.wp[0][H * kHalf + c];
            a = fmaf(yp, p.w.wp[1][H * kHalf + c], a);
            a = fmaf(zp, p.w.wp[2][H * kHalf + c], a);
            a = fmaf(in, p.w.wp[3][H * kHalf + c], a);
            a = fmaf(tm, p.w.wp[4][H * kHalf + c], a);
            col[c * kLd] = a;
        }
    } else {
#pragma unroll
        for (int c = 0; c < kHalf; ++c) col[c * kLd] = -INFINITY;
    }
}

// kFT = threads per CTA == list positions per chunk
template <int kFT>
__global__ void __launch_bounds__(kFT, 1024 / kFT) k_pillar_features_fast(const __grid_constant__ FastParams p)
{
    constexpr int kLd = kFT + kTailW + 1;
    extern __shared__ __align__(16) float s_y[];  // [kHalf][kLd]
    __shared__ float4 s_c4[kFT];  // pillar centre x,y,z ; w = 1.0 when the pillar has empty (padded) slots
    __shared__ float4 s_m4[kFT];  // mean - centre x,y,z ; w = output row as int bits (-1: pillar not emitted)
    __shared__ uint32_t s_thr[kFT], s_pos0[kFT], s_cnt[kFT];
    __shared__ float s_px[kFT], s_py[kFT], s_pz[kFT];  // the chunk's coordinates, for the per-pillar mean
    __shared__ int16_t s_start[kFT + 1];
    __shared__ uint16_t s_big[kFT];
    __shared__ int s_nbig, s_tail_len;
    __shared__ int s_warp_cnt[kFT / 32];
    __shared__ __align__(16) float s_wk[7][64];
    __shared__ float s_acc_tail[kHalf];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t total = p.hdr->total_listed;
    const uint32_t q0 = blockIdx.x * kFT;
    if (q0 >= total) return;

    for (int i = tid; i < 7 * 64; i += kFT) s_wk[i >> 6][i & 63] = (i < 6 * 64) ? p.w.wk[i >> 6][i & 63] : p.w.shift[i & 63];

    const uint32_t pos = q0 + tid;
    const bool in = pos < total;
    float4 ra = make_float4(0.f, 0.f, 0.f, 0.f), rb = make_float4(0.f, 0.f, 0.f, __uint_as_float(1u));
    if (in) {
        const float4 *src = reinterpret_cast<const float4 *>(p.records + pos);
        ra = __ldg(src);
        rb = __ldg(src + 1);
    }
    const uint32_t r_idx = __float_as_uint(rb.y), r_gid = __float_as_uint(rb.z);
    const bool is_start = in && __float_as_uint(rb.w) == 0u;
    s_px[tid] = ra.x;
    s_py[tid] = ra.y;
    s_pz[tid] = ra.z;

    // pillar-local index of every position: inclusive count of list starts, minus one
    const unsigned bal = __ballot_sync(kFull, is_start);
    if (lane == 0) s_warp_cnt[warp] = __popc(bal);
    if (tid == 0) {
        s_nbig = 0;
        s_tail_len = 0;
    }
    __syncthreads();
    int before = 0, n_pl = 0;
#pragma unroll
    for (int w = 0; w < kFT / 32; ++w) {
        const int c = s_warp_cnt[w];
        if (w < warp) before += c;
        n_pl += c;
    }
    if (n_pl == 0) return;  // every position belongs to a pillar owned by an earlier CTA
    const int pl = before + __popc(bal & lanemask_le()) - 1;
    if (is_start) s_start[pl] = static_cast<int16_t>(tid);
    if (tid == 0) s_start[n_pl] = kFT;

    // ---- per-pillar constants, by the thread sitting on the list start ---------------------------------------------
    const int P = p.gd.max_points;
    if (is_start) {
        const uint32_t key = __ldg(p.pillar_key + r_gid);
        const uint32_t n = __ldg(p.pillar_cnt + r_gid);
        const uint32_t b = key / p.gd.cells;
        const uint32_t cell = key - b * p.gd.cells;
        const uint32_t z = cell / p.gd.cells_xy;
        const uint32_t rem = cell - z * p.gd.cells_xy;
        const uint32_t y = rem / static_cast<uint32_t>(p.gd.g[0]);
        const uint32_t x = rem - y * static_cast<uint32_t>(p.gd.g[0]);
        const uint32_t local = r_gid - __ldg(p.frame_gstart + b);
        const int64_t row = static_cast<int64_t>(__ldg(p.frame_rowbase + b)) + local;
        const bool live = local < static_cast<uint32_t>(p.gd.max_voxels) && row < p.capacity;
        // pillar centre: coord * voxel + offset, two roundings as in the reference (pillar_vfe.py:101-103)
        const float cx = __fadd_rn(__fmul_rn(static_cast<float>(x), p.vsz[0]), p.off[0]);
        const float cy = __fadd_rn(__fmul_rn(static_cast<float>(y), p.vsz[1]), p.off[1]);
        const float cz = __fadd_rn(__fmul_rn(static_cast<float>(z), p.vsz[2]), p.off[2]);
        s_c4[pl] = make_float4(cx, cy, cz, n < static_cast<uint32_t>(P) ? 1.f : 0.f);
        float4 m4 = make_float4(0.f, 0.f, 0.f, __int_as_float(live ? static_cast<int32_t>(row) : -1));
        s_cnt[pl] = n;
        s_pos0[pl] = pos;
        s_thr[pl] = 0xFFFFFFFFu;
        if (live) {
            if (n > static_cast<uint32_t>(P)) {
                s_big[atomicAdd(&s_nbig, 1)] = static_cast<uint16_t>(pl);
            } else {
                double sx = 0.0, sy = 0.0, sz = 0.0;  // double: the sum does not depend on the list order
                for (uint32_t j = 0; j < n; ++j) {
                    const uint32_t t = tid + j;
                    if (t < kFT) {  // staged by the owner of that position before the first barrier
                        sx += static_cast<double>(s_px[t]);
                        sy += static_cast<double>(s_py[t]);
                        sz += static_cast<double>(s_pz[t]);
                    } else {  // the pillar continues in the next chunk
                        const float4 q = __ldg(reinterpret_cast<const float4 *>(p.records + pos + j));
                        sx += static_cast<double>(q.x);
                        sy += static_cast<double>(q.y);
                        sz += static_cast<double>(q.z);
                    }
                }
                const float nf = static_cast<float>(n);
                m4.x = __fsub_rn(__fdiv_rn(static_cast<float>(sx), nf), cx);  // pillar_vfe.py:97, relative to the centre
                m4.y = __fsub_rn(__fdiv_rn(static_cast<float>(sy), nf), cy);
                m4.z = __fsub_rn(__fdiv_rn(static_cast<float>(sz), nf), cz);
            }
            if (p.voxel_coords)
                *reinterpret_cast<int4 *>(p.voxel_coords + row * 4) =
                    make_int4(static_cast<int>(b), static_cast<int>(z), static_cast<int>(y), static_cast<int>(x));
            if (p.voxel_num_points) p.voxel_num_points[row] = static_cast<int32_t>(min(n, static_cast<uint32_t>(P)));
            if (p.cell_row)
                p.cell_row[static_cast<int64_t>(b) * p.gd.cells_xy + static_cast<int64_t>(y) * p.gd.g[0] + x] =
                    static_cast<int32_t>(row);
            if (pl == n_pl - 1) {
                const int64_t over = static_cast<int64_t>(pos) + n - (static_cast<int64_t>(q0) + kFT);
                s_tail_len = over > 0 ? static_cast<int>(over) : 0;
            }
        }
        s_m4[pl] = m4;
    }
    __syncthreads();

    // ---- pillars over the cap: threshold = P-th smallest point index (radix select), mean over the kept ones ---------
    for (int k = warp; k < s_nbig; k += kFT / 32) {
        const int bp = s_big[k];
        const uint32_t p0 = s_pos0[bp], n = s_cnt[bp];
        uint32_t prefix = 0, kk = static_cast<uint32_t>(P);
        for (int bit = p.idx_bits - 1; bit >= 0; --bit) {
            const uint32_t himask = 0xFFFFFFFFu << (bit + 1);
            uint32_t c0 = 0;
            for (uint32_t j = lane; j < n; j += 32) {
                const uint32_t v = __ldg(&p.records[p0 + j].idx);
                c0 += ((v & himask) == prefix && ((v >> bit) & 1u) == 0u) ? 1u : 0u;
            }
#pragma unroll
            for (int s = 16; s > 0; s >>= 1) c0 += __shfl_xor_sync(kFull, c0, s);
            if (kk > c0) {
                prefix |= 1u << bit;
                kk -= c0;
            }
        }
        double sx = 0.0, sy = 0.0, sz = 0.0;
        for (uint32_t j = lane; j < n; j += 32) {
            const float4 q = __ldg(reinterpret_cast<const float4 *>(p.records + p0 + j));
            if (__ldg(&p.records[p0 + j].idx) <= prefix) {
                sx += static_cast<double>(q.x);
                sy += static_cast<double>(q.y);
                sz += static_cast<double>(q.z);
            }
        }
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
            sx += __shfl_xor_sync(kFull, sx, s);
            sy += __shfl_xor_sync(kFull, sy, s);
            sz += __shfl_xor_sync(kFull, sz, s);
        }
        if (lane == 0) {
            const float nf = static_cast<float>(P);
            const float4 c4 = s_c4[bp];
            s_thr[bp] = prefix;
            s_m4[bp].x = __fsub_rn(__fdiv_rn(static_cast<float>(sx), nf), c4.x);
            s_m4[bp].y = __fsub_rn(__fdiv_rn(static_cast<float>(sy), nf), c4.y);
            s_m4[bp].z = __fsub_rn(__fdiv_rn(static_cast<float>(sz), nf), c4.z);
        }
    }
    __syncthreads();

    // ---- my point, relative to its pillar centre ---------------------------------------------------------------------
    bool kept = false;
    float xp = 0.f, yp = 0.f, zp = 0.f;
    if (in && pl >= 0 && __float_as_int(s_m4[pl].w) >= 0 && r_idx <= s_thr[pl]) {
        kept = true;
        const float4 c4 = s_c4[pl];
        xp = __fsub_rn(ra.x, c4.x);
        yp = __fsub_rn(ra.y, c4.y);
        zp = __fsub_rn(ra.z, c4.z);
    }
    const int tail_len = s_tail_len;
    const int last = n_pl - 1;
    const int sub = tid & 7, slot = tid >> 3;  // max stage: 8 threads x 4 channels per pillar, 32 pillars per sweep

#pragma unroll
    for (int H = 0; H < 2; ++H) {
        if (H == 0) point_half<0, kLd>(p, kept, xp, yp, zp, ra.w, rb.x, s_y + tid);
        else point_half<1, kLd>(p, kept, xp, yp, zp, ra.w, rb.x, s_y + tid);
        if (tid < kHalf) s_acc_tail[tid] = -INFINITY;
        __syncthreads();

        // tail of the last pillar that spills into the following chunk(s)
        for (int t0 = 0; t0 < tail_len; t0 += kTailW) {
            if (warp == 0) {
                const bool tin = t0 + lane < tail_len;
                bool tk = false;
                float tx = 0.f, ty = 0.f, tz = 0.f, ti = 0.f, tt = 0.f;
                if (tin) {
                    const float4 *src = reinterpret_cast<const float4 *>(p.records + q0 + kFT + t0 + lane);
                    const float4 a = __ldg(src), b = __ldg(src + 1);
                    if (__float_as_uint(b.y) <= s_thr[last]) {
                        const float4 c4 = s_c4[last];
                        tk = true;
                        tx = __fsub_rn(a.x, c4.x);
                        ty = __fsub_rn(a.y, c4.y);
                        tz = __fsub_rn(a.z, c4.z);
                        ti = a.w;
                        tt = b.x;
                    }
                }
                if (H == 0) point_half<0, kLd>(p, tk, tx, ty, tz, ti, tt, s_y + kFT + lane);
                else point_half<1, kLd>(p, tk, tx, ty, tz, ti, tt, s_y + kFT + lane);
            }
            __syncthreads();
            if (tid < kHalf) {
                float m = s_acc_tail[tid];
#pragma unroll 8
                for (int l = 0; l < kTailW; ++l) m = fmaxf(m, s_y[tid * kLd + kFT + l]);
                s_acc_tail[tid] = m;
            }
            __syncthreads();
        }

        // max over each pillar's points, then the per-pillar constant, ReLU and the padded-slot term
        {
            const int ch = H * kHalf + 4 * sub;
            const float4 k0 = *reinterpret_cast<const float4 *>(&s_wk[0][ch]), k1 = *reinterpret_cast<const float4 *>(&s_wk[1][ch]),
                         k2 = *reinterpret_cast<const float4 *>(&s_wk[2][ch]), k3 = *reinterpret_cast<const float4 *>(&s_wk[3][ch]),
                         k4 = *reinterpret_cast<const float4 *>(&s_wk[4][ch]), k5 = *reinterpret_cast<const float4 *>(&s_wk[5][ch]),
                         sh = *reinterpret_cast<const float4 *>(&s_wk[6][ch]);
            const float4 relu_sh = make_float4(fmaxf(sh.x, 0.f), fmaxf(sh.y, 0.f), fmaxf(sh.z, 0.f), fmaxf(sh.w, 0.f));
            const float *ycol = s_y + (4 * sub) * kLd;
            for (int q = slot; q < n_pl; q += kFT / 8) {
                const float4 m4 = s_m4[q];
                const int row = __float_as_int(m4.w);
                if (row < 0) continue;
                const float4 c4 = s_c4[q];
                const int a0 = s_start[q], a1 = s_start[q + 1];
                // almost every pillar has <= 4 points in the chunk: four clamped, branch-free reads (re-reading the last
                // element is harmless for a max), then a loop only for the rare longer run
                const int l = a1 - 1;
                const int j1 = min(a0 + 1, l), j2 = min(a0 + 2, l), j3 = min(a0 + 3, l);
                float m0 = fmaxf(fmaxf(ycol[a0], ycol[j1]), fmaxf(ycol[j2], ycol[j3]));
                float m1 = fmaxf(fmaxf(ycol[kLd + a0], ycol[kLd + j1]), fmaxf(ycol[kLd + j2], ycol[kLd + j3]));
                float m2 = fmaxf(fmaxf(ycol[2 * kLd + a0], ycol[2 * kLd + j1]), fmaxf(ycol[2 * kLd + j2], ycol[2 * kLd + j3]));
                float m3 = fmaxf(fmaxf(ycol[3 * kLd + a0], ycol[3 * kLd + j1]), fmaxf(ycol[3 * kLd + j2], ycol[3 * kLd + j3]));
                for (int j = a0 + 4; j < a1; ++j) {
                    m0 = fmaxf(m0, ycol[j]);
                    m1 = fmaxf(m1, ycol[kLd + j]);
                    m2 = fmaxf(m2, ycol[2 * kLd + j]);
                    m3 = fmaxf(m3, ycol[3 * kLd + j]);
                }
                if (q == last && tail_len > 0) {
                    m0 = fmaxf(m0, s_acc_tail[4 * sub + 0]);
                    m1 = fmaxf(m1, s_acc_tail[4 * sub + 1]);
                    m2 = fmaxf(m2, s_acc_tail[4 * sub + 2]);
                    m3 = fmaxf(m3, s_acc_tail[4 * sub + 3]);
                }
                const bool pad = c4.w != 0.f;
                float4 out;
                {
                    float kc = fmaf(k0.x, c4.x, sh.x); kc = fmaf(k1.x, c4.y, kc); kc = fmaf(k2.x, c4.z, kc);
                    kc = fmaf(-k3.x, m4.x, kc); kc = fmaf(-k4.x, m4.y, kc); kc = fmaf(-k5.x, m4.z, kc);
                    out.x = fmaxf(m0 + kc, pad ? relu_sh.x : 0.f);
                }
                {
                    float kc = fmaf(k0.y, c4.x, sh.y); kc = fmaf(k1.y, c4.y, kc); kc = fmaf(k2.y, c4.z, kc);
                    kc = fmaf(-k3.y, m4.x, kc); kc = fmaf(-k4.y, m4.y, kc); kc = fmaf(-k5.y, m4.z, kc);
                    out.y = fmaxf(m1 + kc, pad ? relu_sh.y : 0.f);
                }
                {
                    float kc = fmaf(k0.z, c4.x, sh.z); kc = fmaf(k1.z, c4.y, kc); kc = fmaf(k2.z, c4.z, kc);
                    kc = fmaf(-k3.z, m4.x, kc); kc = fmaf(-k4.z, m4.y, kc); kc = fmaf(-k5.z, m4.z, kc);
                    out.z = fmaxf(m2 + kc, pad ? relu_sh.z : 0.f);
                }
                {
                    float kc = fmaf(k0.w, c4.x, sh.w); kc = fmaf(k1.w, c4.y, kc); kc = fmaf(k2.w, c4.z, kc);
                    kc = fmaf(-k3.w, m4.x, kc); kc = fmaf(-k4.w, m4.y, kc); kc = fmaf(-k5.w, m4.z, kc);
                    out.w = fmaxf(m3 + kc, pad ? relu_sh.w : 0.f);
                }
                *reinterpret_cast<float4 *>(p.pillar_features + static_cast<int64_t>(row) * 64 + ch) = out;
            }
        }
        if (H == 0) __syncthreads();  // the second half overwrites s_y
    }
}

}  // namespace

cudaError_t launch_pillar_features_fast(const FastJob &job, const FastWeights &w, const GridDev &gd, const Workspace &ws,
                                        cudaStream_t st)
{
    if (job.n == 0) return cudaSuccess;
    FastParams p{};
    p.records = ws.records;
    p.hdr = ws.hdr;
    p.pillar_key = ws.pillar_key;
    p.pillar_cnt = ws.pillar_cnt;
    p.frame_gstart = ws.frame_gstart;
    p.frame_rowbase = ws.frame_rowbase;
    p.cell_row = job.write_cell_row ? ws.cell_row : nullptr;
    p.pillar_features = job.pillar_features;
    p.voxel_coords = job.voxel_coords;
    p.voxel_num_points = job.voxel_num_points;
    p.capacity = job.capacity;
    p.gd = gd;
    for (int i = 0; i < 3; ++i) {
        p.vsz[i] = job.vsz[i];
        p.off[i] = job.off[i];
    }
    p.idx_bits = job.idx_bits;
    p.w = w;
    static int ft = 0;
    if (!ft) {
        const char *e = getenv("PILLARS_FEAT_THREADS");
        ft = e ? atoi(e) : 256;
        if (ft != 64 && ft != 128 && ft != 256) ft = 256;
    }
    const size_t smem = sizeof(float) * kHalf * (ft + kTailW + 1);
    const unsigned grid = static_cast<unsigned>((job.n + ft - 1) / ft);  // upper bound: listed points <= n
    // static + dynamic shared memory is just over the 48 KB default at 256 threads
    if (ft == 256) {
        cudaFuncSetAttribute(k_pillar_features_fast<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        k_pillar_features_fast<256><<<grid, 256, smem, st>>>(p);
    } else if (ft == 128) {
        k_pillar_features_fast<128><<<grid, 128, smem, st>>>(p);
    } else {
        k_pillar_features_fast<64><<<grid, 64, smem, st>>>(p);
    }
    note_launch();
    return cudaGetLastError();
}

}  // namespace pillars
