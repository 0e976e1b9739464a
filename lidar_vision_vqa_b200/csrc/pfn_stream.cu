// Per-pillar feature stage for the mainstream configuration (USE_ABSLOTE_XYZ, no WITH_DISTANCE, C <= 5 point channels,
// one PFN layer, F = 64): pillar_vfe.py:94-123 (augment) + :29-49 (Linear + BatchNorm(eval) + ReLU + max over the pillar).
//
// Shape of the work.  The points arrive as 32-byte records grouped by pillar (k_place), ~2 points per pillar.  The layer is
// regrouped around the pillar centre c (x' = x - c, m' = mean - c):
//       W.[p, p_xyz - mean, p_xyz - c] = (W_p + W_cl + W_ce).x' + W_it.(i,t)  +  [ W_p.c - W_cl.m' ]
// i.e. 5 FMAs per (point, channel) plus one 6-FMA constant per (pillar, channel); every per-point term is a small number, so
// nothing cancels.  BatchNorm's scale is folded into W, and since "+ constant" and ReLU are monotone the max over the
// pillar's points is taken before them.
//
// Mapping.  The kernel is bound by instruction issue, not by HBM, so the layout minimises issued instructions per point:
//   * phase 1, one thread per list position (CTA = kFT positions + a 32-position look-ahead): stage the records in shared
//     memory; the thread sitting on a list start derives the pillar's row, centre, mean and writes voxel_coords /
//     voxel_num_points / the BEV index map; every thread then rewrites its point relative to the centre (NaN when the point
//     is beyond the first-P cap: fmaxf ignores NaN, so dropped points need no branch later);
//   * phase 2, one warp per run of pillars, ONE LANE PER CHANNEL PAIR: the warp walks its points in list order, each point
//     is two broadcast shared-memory loads + 5 packed FFMA2 (fma.rn.f32x2: two channels per instruction, the point value
//     as the scalar-broadcast operand) + 2 FMNMX with the running max in registers; at a pillar end (a flag stored with
//     the point) the lanes add the per-pillar constant (6 FFMA2), apply ReLU / the padded-slot term and store the 256-byte
//     output row with one coalesced 8-byte store per lane.  No shared-memory transpose, no divergence, no idle lanes on
//     short pillars.
// A CTA owns the pillars whose list STARTS inside its kFT positions; a list that runs past the look-ahead (only possible
// for n > 32) is finished from global memory by the owning warp.  Pillars over the cap (n > P) get their "first P by point
// index" threshold from a warp-wide radix select.
#include <cstdlib>

#include "common.cuh"

namespace pillars {

namespace {

constexpr int kLook = 32;  // list positions staged beyond the CTA's own chunk
constexpr unsigned kFull = 0xffffffffu;
constexpr int kFlagLast = 1;  // final point of its pillar
constexpr int kFlagStop = 2;  // ... and that pillar is the last one owned by the walking warp
constexpr int kFlagMore = 4;  // the pillar continues beyond the staged positions

struct StreamParams {
    const PointRecord *records;   // grouped by pillar, x,y,z relative to the pillar centre (k_place)
    const Header *hdr;
    const float4 *pillar_meta;    // [2 * pillars] by pillar id, see Workspace
    const float *folded;          // [PILLARS_FOLDED_FLOATS], see launch_fold_pfn
    float *pillar_features;
    int max_points;
    int idx_bits;
};

// ---- packed fp32 pairs (Blackwell FFMA2 / FMUL2 / FADD2) ------------------------------------------------------------
__device__ __forceinline__ unsigned long long pack2(float2 v)
{
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(v.x), "f"(v.y));
    return r;
}
__device__ __forceinline__ float2 unpack2(unsigned long long r)
{
    float2 v;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(v.x), "=f"(v.y) : "l"(r));
    return v;
}
// w * (x, x) + c : ptxas turns the duplicated half into FFMA2's scalar-broadcast operand
__device__ __forceinline__ float2 fma2s(float2 w, float x, float2 c)
{
    unsigned long long rx, rd;
    asm("mov.b64 %0, {%1, %1};" : "=l"(rx) : "f"(x));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(pack2(w)), "l"(rx), "l"(pack2(c)));
    return unpack2(rd);
}
__device__ __forceinline__ float2 mul2s(float2 w, float x)
{
    unsigned long long rx, rd;
    asm("mov.b64 %0, {%1, %1};" : "=l"(rx) : "f"(x));
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(pack2(w)), "l"(rx));
    return unpack2(rd);
}
__device__ __forceinline__ float2 add2(float2 a, float2 b)
{
    unsigned long long rd;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(pack2(a)), "l"(pack2(b)));
    return unpack2(rd);
}

// shared-memory accesses of the walk: 32-bit addresses, so that the loop carries one address register and no generic-pointer
// arithmetic
__device__ __forceinline__ float4 lds4(uint32_t addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ float2 lds2(uint32_t addr)
{
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
    return v;
}

struct LaneWeights {
    float2 w0, w1, w2, w3, w4;        // per point: x', y', z', intensity, time
    float2 k0, k1, k2, k3, k4, k5;    // per pillar: centre xyz, -(cluster weights) on (mean - centre)
    float2 sh, rsh;                   // BatchNorm shift, relu(shift)
};

__device__ __forceinline__ void point_step(const LaneWeights &w, const float4 a, const float t, float2 &acc)
{
    float2 y = mul2s(w.w0, a.x);
    y = fma2s(w.w1, a.y, y);
    y = fma2s(w.w2, a.z, y);
    y = fma2s(w.w3, a.w, y);
    y = fma2s(w.w4, t, y);
    acc.x = fmaxf(acc.x, y.x);  // a NaN y (dropped point) leaves acc unchanged
    acc.y = fmaxf(acc.y, y.y);
}

// One warp = 32 consecutive list positions (+ a 32-position look-ahead), nothing shared between warps: no CTA barrier, a
// warp that is done leaves.  kWarps warps per CTA only share the launch.
template <int kWarps>
__global__ void __launch_bounds__(32 * kWarps, 2048 / (32 * kWarps) > 32 ? 32 : 2048 / (32 * kWarps) / 2)
k_pillar_features_stream(const __grid_constant__ StreamParams p)
{
    constexpr int kOwn = 32;
    constexpr int kStage = kOwn + kLook;
    // Per warp, one 32-byte slot per staged list position for the point ("pt") and one per own position for the pillar that
    // starts there ("pl"); the walk addresses both from one running shared-memory address.
    //   pt[2j]   = x, y, z (relative to the pillar centre), intensity ;  x = NaN: point beyond the first-P cap
    //   pt[2j+1] = time, flags (int bits), point index, position inside the pillar's list
    //   pl[2j]   = centre x,y,z ; w = 1.0 when the pillar has empty (padded) slots
    //   pl[2j+1] = mean - centre x,y,z ; w = output row as int bits (-1: pillar not emitted)
    constexpr int kPl = 2 * (kStage + 1);       // float4 offset of the pl half
    constexpr int kSlot = kPl + 2 * kOwn;       // float4 per warp
    __shared__ float4 s_all[kWarps * kSlot];
    __shared__ uint32_t s_thr_all[kWarps * kOwn];  // largest kept point index of the pillar starting there

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t total = p.hdr->total_listed;
    const uint32_t q0 = (blockIdx.x * kWarps + warp) * kOwn;
    if (q0 >= total) return;
    float4 *const s_pt = s_all + warp * kSlot;
    float4 *const s_pl = s_pt + kPl;
    uint32_t *const s_thr = s_thr_all + warp * kOwn;
    const float qnan = __int_as_float(0x7fc00000);

    // ---- phase 1a: stage 64 records ---------------------------------------------------------------------------------------
    const uint32_t pos = q0 + lane;
    float4 ra = make_float4(qnan, 0.f, 0.f, 0.f), rb = make_float4(0.f, 0.f, 0.f, __uint_as_float(1u));
    float4 ta = ra, tb = make_float4(0.f, 0.f, 0.f, __uint_as_float(0xFFFFFFFFu));
    if (pos < total) {
        const float4 *src = reinterpret_cast<const float4 *>(p.records + pos);
        ra = __ldg(src);
        rb = __ldg(src + 1);
    }
    if (pos + kOwn < total) {
        const float4 *src = reinterpret_cast<const float4 *>(p.records + pos + kOwn);
        ta = __ldg(src);
        tb = __ldg(src + 1);
    }
    const uint32_t r_arr = __float_as_uint(rb.w);
    const bool is_start = r_arr == 0u;  // positions beyond the list carry arrival 1
    const unsigned bal = __ballot_sync(kFull, is_start);
    if (bal == 0u) return;  // every position belongs to a pillar that started in an earlier chunk
    float4 m0 = make_float4(0.f, 0.f, 0.f, 0.f), m1 = m0;
    if (is_start) {
        m0 = __ldg(p.pillar_meta + 2 * static_cast<size_t>(pos));
        m1 = __ldg(p.pillar_meta + 2 * static_cast<size_t>(pos) + 1);
    }
    s_pt[2 * lane] = ra;
    s_pt[2 * lane + 1] = rb;
    s_pt[2 * (kOwn + lane)] = ta;
    s_pt[2 * (kOwn + lane) + 1] = tb;
    if (lane == 0) {
        s_pt[2 * kStage] = make_float4(qnan, 0.f, 0.f, 0.f);
        s_pt[2 * kStage + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncwarp();

    // ---- phase 1b: the lane sitting on a list start publishes the pillar's constants and marks its final point -------------
    const uint32_t P = static_cast<uint32_t>(p.max_points);
    const uint32_t n = __float_as_uint(m1.y);
    const bool live = is_start && __float_as_int(m1.x) >= 0;
    const bool big = live && n > P;
    int more = 0;  // points of the warp's last pillar beyond the staged positions
    if (is_start) {
        float4 m4 = make_float4(0.f, 0.f, 0.f, m1.x);
        s_thr[lane] = 0xFFFFFFFFu;
        if (live && !big) {
            // mean of the pillar's points (pillar_vfe.py:97); double: the sum does not depend on the list order
            double sx = 0.0, sy = 0.0, sz = 0.0;
            const uint32_t n_in = min(n, static_cast<uint32_t>(kStage - lane));
            for (uint32_t j = 0; j < n_in; ++j) {
                const float4 q = s_pt[2 * (lane + j)];
                sx += static_cast<double>(q.x);
                sy += static_cast<double>(q.y);
                sz += static_cast<double>(q.z);
            }
            for (uint32_t j = n_in; j < n; ++j) {  // P > 32 only
                const float4 q = __ldg(reinterpret_cast<const float4 *>(p.records + pos + j));
                sx += static_cast<double>(q.x);
                sy += static_cast<double>(q.y);
                sz += static_cast<double>(q.z);
            }
            // the reference rounds the ABSOLUTE mean to fp32 before subtracting it; reproduce that rounding step
            const float rn = __frcp_rn(static_cast<float>(n));
            m4.x = __fsub_rn(__fadd_rn(m0.x, static_cast<float>(sx) * rn), m0.x);
            m4.y = __fsub_rn(__fadd_rn(m0.y, static_cast<float>(sy) * rn), m0.y);
            m4.z = __fsub_rn(__fadd_rn(m0.z, static_cast<float>(sz) * rn), m0.z);
        }
        s_pl[2 * lane] = m0;
        s_pl[2 * lane + 1] = m4;
        // walk control: flag the pillar's final point (dropped pillars are walked too, their row is -1)
        const uint32_t endp = static_cast<uint32_t>(lane) + n - 1u;
        const bool last_of_warp = (31 - __clz(bal)) == lane;
        if (endp < static_cast<uint32_t>(kStage)) {
            s_pt[2 * endp + 1].y = __int_as_float(last_of_warp ? kFlagStop : kFlagLast);
        } else {  // only the warp's last pillar can run past the look-ahead
            s_pt[2 * (kStage - 1) + 1].y = __int_as_float(kFlagStop | kFlagMore);
            more = static_cast<int>(endp + 1u - kStage);
        }
    }
    more = __shfl_sync(kFull, more, 31 - __clz(bal));
    __syncwarp();

    // ---- pillars over the cap: threshold = P-th smallest point index (radix select), mean over the kept ones ---------
    unsigned bigmask = __ballot_sync(kFull, big);
    if (bigmask) {
        while (bigmask) {
            const int bp = __ffs(bigmask) - 1;
            bigmask &= bigmask - 1;
            const uint32_t p0 = q0 + bp, nb = __shfl_sync(kFull, n, bp);
            uint32_t prefix = 0, kk = P;
            for (int bit = p.idx_bits - 1; bit >= 0; --bit) {
                const uint32_t himask = 0xFFFFFFFFu << (bit + 1);
                uint32_t c0 = 0;
                for (uint32_t j = lane; j < nb; j += 32) {
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
            for (uint32_t j = lane; j < nb; j += 32) {
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
                const float rn = __frcp_rn(static_cast<float>(P));
                const float4 c4 = s_pl[2 * bp];
                s_thr[bp] = prefix;
                s_pl[2 * bp + 1].x = __fsub_rn(__fadd_rn(c4.x, static_cast<float>(sx) * rn), c4.x);
                s_pl[2 * bp + 1].y = __fsub_rn(__fadd_rn(c4.y, static_cast<float>(sy) * rn), c4.y);
                s_pl[2 * bp + 1].z = __fsub_rn(__fadd_rn(c4.z, static_cast<float>(sz) * rn), c4.z);
            }
        }
        __syncwarp();
        // points beyond the cap become NaN: fmaxf ignores them, so the walk needs no branch
        if (r_arr <= static_cast<uint32_t>(lane) && __float_as_uint(rb.z) > s_thr[lane - static_cast<int>(r_arr)])
            s_pt[2 * lane].x = qnan;
        {
            const uint32_t t_arr = __float_as_uint(tb.w);
            const int j = kOwn + lane;
            const int ps = j - static_cast<int>(t_arr);
            if (t_arr <= static_cast<uint32_t>(j) && ps < kOwn && __float_as_uint(tb.z) > s_thr[ps]) s_pt[2 * j].x = qnan;
        }
        __syncwarp();
    }

    // ---- phase 2: the warp walks the pillars that start among its 32 positions; lane = channel pair -------------------------
    LaneWeights w;
    {
        const float2 *fw = reinterpret_cast<const float2 *>(p.folded) + lane;
        w.w0 = __ldg(fw + 0 * 32); w.w1 = __ldg(fw + 1 * 32); w.w2 = __ldg(fw + 2 * 32); w.w3 = __ldg(fw + 3 * 32);
        w.w4 = __ldg(fw + 4 * 32);
        w.k0 = __ldg(fw + 5 * 32); w.k1 = __ldg(fw + 6 * 32); w.k2 = __ldg(fw + 7 * 32); w.k3 = __ldg(fw + 8 * 32);
        w.k4 = __ldg(fw + 9 * 32); w.k5 = __ldg(fw + 10 * 32);
        w.sh = __ldg(fw + 11 * 32); w.rsh = __ldg(fw + 12 * 32);
    }
    // the lane's output base as a global-space address kept in registers (not recomputed per pillar)
    unsigned long long out_lane = static_cast<unsigned long long>(__cvta_generic_to_global(p.pillar_features + 2 * lane));
    asm volatile("" : "+l"(out_lane));
    constexpr uint32_t kPlBytes = kPl * sizeof(float4);
    const uint32_t s_base = static_cast<uint32_t>(__cvta_generic_to_shared(s_pt));
    uint32_t sp = s_base + 32u * (__ffs(bal) - 1);  // the point being accumulated
    uint32_t ss = sp;  // first point of the pillar being accumulated; its constants sit kPlBytes further on
    float2 acc = make_float2(-INFINITY, -INFINITY);

    // pillar end: per-pillar constant, ReLU, padded-slot term, one 256-byte row
    auto pillar_end = [&]() {
        const float4 c4 = lds4(ss + kPlBytes), m4 = lds4(ss + kPlBytes + 16);
        const int row = __float_as_int(m4.w);
        float2 kc = fma2s(w.k0, c4.x, w.sh);
        kc = fma2s(w.k1, c4.y, kc);
        kc = fma2s(w.k2, c4.z, kc);
        float2 kd = mul2s(w.k3, m4.x);
        kd = fma2s(w.k4, m4.y, kd);
        kd = fma2s(w.k5, m4.z, kd);
        const float2 v = add2(add2(acc, kc), kd);
        const float2 fl2 = mul2s(w.rsh, c4.w);  // relu(shift) when the pillar has padded slots, else 0
        if (row >= 0)
            asm volatile("st.global.v2.f32 [%0], {%1, %2};" ::"l"(out_lane + static_cast<unsigned long long>(static_cast<uint32_t>(row)) * 256ull),
                         "f"(fmaxf(v.x, fl2.x)), "f"(fmaxf(v.y, fl2.y))
                         : "memory");
        acc = make_float2(-INFINITY, -INFINITY);
    };

    // two points per trip, registers ping-ponged so that the next point is always in flight and nothing is copied
    float4 a0 = lds4(sp);
    float2 b0 = lds2(sp + 16);
    int fl;
    while (true) {
        const float4 a1 = lds4(sp + 32);
        const float2 b1 = lds2(sp + 48);
        point_step(w, a0, b0.x, acc);
        fl = __float_as_int(b0.y);
        if (fl != 0) {
            if (fl & kFlagStop) break;
            pillar_end();
            ss = sp + 32;
        }
        a0 = lds4(sp + 64);
        b0 = lds2(sp + 80);
        point_step(w, a1, b1.x, acc);
        fl = __float_as_int(b1.y);
        if (fl != 0) {
            if (fl & kFlagStop) break;
            pillar_end();
            ss = sp + 64;
        }
        sp += 64;
    }
    if (fl & kFlagMore) {
        // rest of a long list, straight from global memory: 32 records per sweep, broadcast by shuffles
        const uint32_t base = q0 + kStage;
        const uint32_t thr = s_thr[(ss - s_base) >> 5];
        for (int k0 = 0; k0 < more; k0 += 32) {
            const int cnt = min(32, more - k0);
            float4 qa = make_float4(qnan, 0.f, 0.f, 0.f);
            float qt = 0.f;
            if (lane < cnt) {
                const float4 *src = reinterpret_cast<const float4 *>(p.records + base + k0 + lane);
                const float4 u = __ldg(src), v = __ldg(src + 1);
                qa = u;
                if (__float_as_uint(v.z) > thr) qa.x = qnan;
                qt = v.x;
            }
            for (int l = 0; l < cnt; ++l) {
                const float4 v = make_float4(__shfl_sync(kFull, qa.x, l), __shfl_sync(kFull, qa.y, l),
                                             __shfl_sync(kFull, qa.z, l), __shfl_sync(kFull, qa.w, l));
                point_step(w, v, __shfl_sync(kFull, qt, l), acc);
            }
        }
    }
    pillar_end();
}

// ---- folding of the layer's weights (device side, once per call inside k_place's idle lanes would also do; kept as its
//      own tiny kernel so that callers can prepare the table once per model) ------------------------------------------
__global__ void k_fold_pfn(const float *__restrict__ weight, const float *__restrict__ scale,
                           const float *__restrict__ shift, int c_point, int c_in, float *__restrict__ folded)
{
    const int o = threadIdx.x;
    if (o >= 64) return;
    const float *w = weight + static_cast<size_t>(o) * c_in;  // feature order: p[0..c), cluster xyz, centre xyz
    const double s = scale[o];
    const int c = c_point;
    for (int a = 0; a < 3; ++a) {
        folded[a * 64 + o] = static_cast<float>(s * (static_cast<double>(w[a]) + w[c + a] + w[c + 3 + a]));
        folded[(5 + a) * 64 + o] = static_cast<float>(s * w[a]);
        folded[(8 + a) * 64 + o] = -static_cast<float>(s * w[c + a]);
    }
    folded[3 * 64 + o] = c > 3 ? static_cast<float>(s * w[3]) : 0.f;
    folded[4 * 64 + o] = c > 4 ? static_cast<float>(s * w[4]) : 0.f;
    const float sh = shift[o];
    folded[11 * 64 + o] = sh;
    folded[12 * 64 + o] = fmaxf(sh, 0.f);
}

}  // namespace

cudaError_t launch_fold_pfn(const PfnDev &pfn, int c_point, int c_in, float *folded, cudaStream_t st)
{
    k_fold_pfn<<<1, 64, 0, st>>>(pfn.weight, pfn.scale, pfn.shift, c_point, c_in, folded);
    note_launch();
    return cudaGetLastError();
}

cudaError_t launch_pillar_features_stream(const FastJob &job, const float *folded, const GridDev &gd, const Workspace &ws,
                                          cudaStream_t st)
{
    if (job.n == 0) return cudaSuccess;
    StreamParams p{};
    p.records = ws.records;
    p.hdr = ws.hdr;
    p.pillar_meta = ws.pillar_meta;
    p.folded = folded;
    p.pillar_features = job.pillar_features;
    p.max_points = gd.max_points;
    p.idx_bits = job.idx_bits;
    static int ft = 0;
    if (!ft) {
        const char *e = getenv("PILLARS_FEAT_THREADS");
        ft = e ? atoi(e) : 128;
        if (ft != 64 && ft != 128 && ft != 256) ft = 128;
    }
    const int64_t chunks = (job.n + 31) / 32;  // upper bound: listed points <= n
    const unsigned grid = static_cast<unsigned>((chunks * 32 + ft - 1) / ft);
    if (ft == 256) k_pillar_features_stream<8><<<grid, 256, 0, st>>>(p);
    else if (ft == 128) k_pillar_features_stream<4><<<grid, 128, 0, st>>>(p);
    else k_pillar_features_stream<2><<<grid, 64, 0, st>>>(p);
    note_launch();
    return cudaGetLastError();
}

}  // namespace pillars
