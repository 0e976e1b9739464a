// Dense BEV canvas in "gather form" (replaces PointPillarScatter.forward,
// src/lidar-encoder/pcdet/models/backbones_2d/map_to_bev/pointpillar_scatter.py:14-37).
//
// The reference zero-fills a [F, ny*nx] canvas per frame, index-assigns F strided 4-byte columns per pillar and then
// stacks the frames (a third full pass).  Here a dense index map cell -> pillar row (4 B per cell, written by the
// feature kernel or by k_build_cell_row) turns the scatter inside out: every output element is written exactly once,
// in NCHW order, with full-line stores, zero-fill included, no atomics.
//
// Three store paths over the same tiling (256 cells x F channels per tile), selectable for measurement:
//   1 plain   : st.global.v4 from registers
//   2 bulk1d  : the tile lives in shared memory ([F][256] floats, all zero except occupied cells); F threads each issue
//               one cp.async.bulk.global.shared::cta row copy (UBLKCP).  An all-empty tile is stored straight from the
//               resident zero tile without touching shared memory.
//   3 tma2d   : same tile, one cp.async.bulk.tensor.2d store per tile through a CUtensorMap over the canvas viewed as
//               [B*F rows, ny*nx cells] (UTMASTG).
#include <cuda.h>
#include <cuda_fp16.h>

#include <cstdlib>

#include "common.cuh"

namespace pillars {

namespace {

constexpr int kThreads = 256;

__global__ void k_build_cell_row(const void *__restrict__ coords, int coords_float, int64_t m,
                                 const int32_t *__restrict__ m_dev, int nb, int nx, int ny, int nz,
                                 int32_t *__restrict__ cell_row)
{
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    int64_t live = m;
    if (m_dev) live = tmin<int64_t>(m, static_cast<int64_t>(*m_dev));
    if (i >= live) return;
    int b, z, y, x;
    if (coords_float) {
        const float4 c = *reinterpret_cast<const float4 *>(static_cast<const float *>(coords) + i * 4);
        b = static_cast<int>(c.x); z = static_cast<int>(c.y); y = static_cast<int>(c.z); x = static_cast<int>(c.w);
    } else {
        const int4 c = *reinterpret_cast<const int4 *>(static_cast<const int32_t *>(coords) + i * 4);
        b = c.x; z = c.y; y = c.z; x = c.w;
    }
    // pointpillar_scatter.py:27  index = z + y*nx + x (nz == 1, so z == 0);  :63  index = z*ny*nx + y*nx + x (3-D variant)
    const int64_t plane = static_cast<int64_t>(nx) * ny * nz;
    const int64_t cell = nz > 1 ? (static_cast<int64_t>(z) * ny + y) * nx + x : static_cast<int64_t>(z) + static_cast<int64_t>(y) * nx + x;
    if (b < 0 || b >= nb || cell < 0 || cell >= plane || x < 0 || x >= nx) return;
    atomicMax(cell_row + static_cast<int64_t>(b) * plane + cell, static_cast<int32_t>(i));  // later row wins
}

// ---- variant 1 ----------------------------------------------------------------------------------
// Warp-autonomous direct stores: a warp owns 128 consecutive cells (4 per lane) x all channels = 32 KB of the canvas and
// never waits for another warp.  A warp whose 128 cells are empty streams zeros; otherwise each lane fetches, 8 channels
// at a time, the feature rows of its occupied cells with two 16-byte loads per row and composes one float4 per channel.
// Eight passes of one L2 round trip each per 32 KB keep the SM far above its share of HBM write bandwidth with a handful
// of resident warps, so the kernel runs at the speed of the write stream.
template <bool VEC, int CPP, bool PIPE>
__global__ void __launch_bounds__(kThreads)
k_scatter_plain(const float *__restrict__ feats, const int32_t *__restrict__ cell_row, int f, int64_t plane,
                int tiles_per_plane, float *__restrict__ bev)
{
    constexpr int kPer = VEC ? 4 : 1;
    const int b = blockIdx.x / tiles_per_plane;
    const int64_t cell0 = static_cast<int64_t>(blockIdx.x % tiles_per_plane) * (blockDim.x * kPer) + threadIdx.x * kPer;
    const bool inb = cell0 < plane;
    int32_t r[kPer];
#pragma unroll
    for (int k = 0; k < kPer; ++k) r[k] = -1;
    if (inb) {
        if (VEC) {
            const int4 v = __ldg(reinterpret_cast<const int4 *>(cell_row + b * plane + cell0));
            r[0] = v.x; r[1 % kPer] = v.y; r[2 % kPer] = v.z; r[3 % kPer] = v.w;
        } else {
            r[0] = __ldg(cell_row + b * plane + cell0);
        }
    }
    bool any = false;
#pragma unroll
    for (int k = 0; k < kPer; ++k) any |= r[k] >= 0;
    float *dst = bev + (static_cast<int64_t>(b) * f) * plane + cell0;
    if (!__any_sync(0xffffffffu, any)) {
        if (inb) {
            if (VEC) {
#pragma unroll 8
                for (int c = 0; c < f; ++c) __stcs(reinterpret_cast<float4 *>(dst + c * plane), make_float4(0.f, 0.f, 0.f, 0.f));
            } else {
                for (int c = 0; c < f; ++c) __stcs(dst + c * plane, 0.f);
            }
        }
        return;
    }
    if (!inb) return;
    if (VEC) {
        constexpr int kV = CPP / 4;  // 16-byte loads per row and pass
        float4 cur[4][kV], nxt[4][kV];
        auto fetch = [&](float4 (&dstv)[4][kV], int c0) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
#pragma unroll
                for (int q = 0; q < kV; ++q) dstv[k][q] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (r[k % kPer] >= 0) {
                    const float4 *row = reinterpret_cast<const float4 *>(feats + static_cast<int64_t>(r[k % kPer]) * f + c0);
#pragma unroll
                    for (int q = 0; q < kV; ++q) dstv[k][q] = __ldg(row + q);
                }
            }
        };
        fetch(cur, 0);
        for (int c0 = 0; c0 < f; c0 += CPP) {
            if (PIPE && c0 + CPP < f) fetch(nxt, c0 + CPP);  // next pass in flight while this one is stored
#pragma unroll
            for (int q = 0; q < kV; ++q) {
                __stcs(reinterpret_cast<float4 *>(dst + (c0 + 4 * q + 0) * plane), make_float4(cur[0][q].x, cur[1][q].x, cur[2][q].x, cur[3][q].x));
                __stcs(reinterpret_cast<float4 *>(dst + (c0 + 4 * q + 1) * plane), make_float4(cur[0][q].y, cur[1][q].y, cur[2][q].y, cur[3][q].y));
                __stcs(reinterpret_cast<float4 *>(dst + (c0 + 4 * q + 2) * plane), make_float4(cur[0][q].z, cur[1][q].z, cur[2][q].z, cur[3][q].z));
                __stcs(reinterpret_cast<float4 *>(dst + (c0 + 4 * q + 3) * plane), make_float4(cur[0][q].w, cur[1][q].w, cur[2][q].w, cur[3][q].w));
            }
            if (c0 + CPP < f) {
                if (PIPE) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
#pragma unroll
                        for (int q = 0; q < kV; ++q) cur[k][q] = nxt[k][q];
                } else {
                    fetch(cur, c0 + CPP);
                }
            }
        }
    } else {
        for (int c = 0; c < f; ++c) __stcs(dst + c * plane, r[0] >= 0 ? __ldg(feats + static_cast<int64_t>(r[0]) * f + c) : 0.f);
    }
}

// ---- variant 4 ----------------------------------------------------------------------------------
// Same warp-autonomous scheme with 256-bit accesses (sm_100: LDG.256 / STG.256): a lane owns 8 consecutive cells, a warp
// 256 cells x all channels = 64 KB of the canvas, and every store instruction of the warp writes 1 KB contiguous.  Half the
// LSU instructions of variant 1 for the same bytes.
__device__ __forceinline__ void stg256(float *p, float a, float b, float c, float d, float e, float f, float g, float h)
{
    asm volatile("st.global.L1::no_allocate.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(a), "f"(b), "f"(c),
                 "f"(d), "f"(e), "f"(f), "f"(g), "f"(h)
                 : "memory");
}
__device__ __forceinline__ void stg256_cs(float *p, float a, float b, float c, float d, float e, float f, float g, float h)
{
    asm volatile("st.global.cs.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d),
                 "f"(e), "f"(f), "f"(g), "f"(h)
                 : "memory");
}

// Work split: a CTA owns blockDim.x * 8 consecutive cells x `chan_per_cta` channels (a multiple of 8).  With 8 channels per
// CTA the canvas is written as few long sequential streams (measured with profiles/micro/fill_patterns.cu on B200: 64
// channel-strided streams per thread 164 us for the 1 GiB canvas, 8 channels x 8 KB runs per CTA 149 us, linear fill 146 us);
// the CTAs of one cell tile then each read one 32-byte sector of every pillar row, so no feature byte is fetched twice.
// tile_major != 0: consecutive CTAs are the channel groups of one cell tile (index map and feature rows are shared by CTAs
// that run together); 0: consecutive CTAs walk the plane for one channel group.
template <bool CS, int CPP>
__global__ void __launch_bounds__(kThreads)
k_scatter_wide(const float *__restrict__ feats, const int32_t *__restrict__ cell_row, int f, int64_t plane,
               int tiles_per_plane, int chan_per_cta, int tile_major, float *__restrict__ bev)
{
    const int groups = f / chan_per_cta;
    int b, tile, cg;
    if (tile_major) {
        cg = blockIdx.x % groups;
        tile = (blockIdx.x / groups) % tiles_per_plane;
        b = blockIdx.x / (groups * tiles_per_plane);
    } else {
        tile = blockIdx.x % tiles_per_plane;
        cg = (blockIdx.x / tiles_per_plane) % groups;
        b = blockIdx.x / (groups * tiles_per_plane);
    }
    const int64_t cell0 = static_cast<int64_t>(tile) * (blockDim.x * 8) + threadIdx.x * 8;
    const bool inb = cell0 < plane;  // plane % 8 == 0: a lane is entirely inside or outside
    int32_t r[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) r[k] = -1;
    if (inb) {
        const int4 *src = reinterpret_cast<const int4 *>(cell_row + b * plane + cell0);
        const int4 v = __ldg(src), w = __ldg(src + 1);
        r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
        r[4] = w.x; r[5] = w.y; r[6] = w.z; r[7] = w.w;
    }
    bool any = false;
#pragma unroll
    for (int k = 0; k < 8; ++k) any |= r[k] >= 0;
    const int c_begin = cg * chan_per_cta, c_end = c_begin + chan_per_cta;
    float *dst = bev + (static_cast<int64_t>(b) * f) * plane + cell0;
    auto store = [&](float *p, float a, float b2, float c, float d, float e, float f2, float g, float h) {
        if (CS) stg256_cs(p, a, b2, c, d, e, f2, g, h);
        else stg256(p, a, b2, c, d, e, f2, g, h);
    };
    if (!__any_sync(0xffffffffu, any)) {
        if (inb) {
#pragma unroll 8
            for (int c = c_begin; c < c_end; ++c) store(dst + c * plane, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f);
        }
        return;
    }
    if (!inb) return;
    constexpr int kV = CPP / 4;  // 16-byte loads per row and pass
    for (int c0 = c_begin; c0 < c_end; c0 += CPP) {
        float4 v[8][kV];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
#pragma unroll
            for (int q = 0; q < kV; ++q) v[k][q] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r[k] >= 0) {
                const float4 *row = reinterpret_cast<const float4 *>(feats + static_cast<int64_t>(r[k]) * f + c0);
#pragma unroll
                for (int q = 0; q < kV; ++q) v[k][q] = __ldg(row + q);
            }
        }
#pragma unroll
        for (int q = 0; q < kV; ++q) {
            float *d = dst + (c0 + 4 * q) * plane;
            store(d, v[0][q].x, v[1][q].x, v[2][q].x, v[3][q].x, v[4][q].x, v[5][q].x, v[6][q].x, v[7][q].x);
            store(d + plane, v[0][q].y, v[1][q].y, v[2][q].y, v[3][q].y, v[4][q].y, v[5][q].y, v[6][q].y, v[7][q].y);
            store(d + 2 * plane, v[0][q].z, v[1][q].z, v[2][q].z, v[3][q].z, v[4][q].z, v[5][q].z, v[6][q].z, v[7][q].z);
            store(d + 3 * plane, v[0][q].w, v[1][q].w, v[2][q].w, v[3][q].w, v[4][q].w, v[5][q].w, v[6][q].w, v[7][q].w);
        }
    }
}

// ---- variant 5 ----------------------------------------------------------------------------------
// Persistent warps over (frame, channel group of 8, 256-cell tile) items, tile fastest.  Per item a warp writes 8 channels x
// 1 KB; the warps of one sweep over the item list cover consecutive tiles, so HBM sees a few long sequential write streams.
// Two latencies are taken off the store path:
//   * the index-map entry of the warp's NEXT item is loaded before the current item is stored;
//   * lanes whose 8 cells are all empty (two thirds of them even in populated areas) store their zeros at once, without
//     waiting for the feature rows the other lanes of the warp gather.
__device__ __forceinline__ void ldg256_stream(const float *p, float (&v)[8])
{
    asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                 : "l"(p));
}

template <bool CS>
__global__ void __launch_bounds__(kThreads)
k_scatter_persist(const float *__restrict__ feats, const int32_t *__restrict__ cell_row, int f, int64_t plane,
                  int tiles_per_plane, int64_t n_items, float *__restrict__ bev)
{
    const int lane = threadIdx.x & 31;
    const int64_t n_warps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    int64_t item = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int groups = f >> 3;
    const int per_frame = groups * tiles_per_plane;

    auto load_index = [&](int64_t it, int4 &lo, int4 &hi) {
        lo = hi = make_int4(-1, -1, -1, -1);
        if (it < n_items) {
            const int b = static_cast<int>(it / per_frame);
            const int tile = static_cast<int>(it % tiles_per_plane);
            const int64_t cell0 = static_cast<int64_t>(tile) * 256 + lane * 8;
            if (cell0 < plane) {
                const int4 *src = reinterpret_cast<const int4 *>(cell_row + b * plane + cell0);
                lo = __ldg(src);
                hi = __ldg(src + 1);
            }
        }
    };
    auto store = [&](float *p, float a, float b2, float c, float d, float e, float f2, float g, float h) {
        if (CS) stg256_cs(p, a, b2, c, d, e, f2, g, h);
        else stg256(p, a, b2, c, d, e, f2, g, h);
    };

    int4 lo, hi;
    load_index(item, lo, hi);
    while (item < n_items) {
        int4 nlo, nhi;
        load_index(item + n_warps, nlo, nhi);

        const int b = static_cast<int>(item / per_frame);
        const int rem = static_cast<int>(item - static_cast<int64_t>(b) * per_frame);
        const int cg = rem / tiles_per_plane;
        const int tile = rem - cg * tiles_per_plane;
        const int64_t cell0 = static_cast<int64_t>(tile) * 256 + lane * 8;
        const int c0 = cg * 8;
        float *dst = bev + (static_cast<int64_t>(b) * f + c0) * plane + cell0;
        if (cell0 < plane) {
            const int32_t r[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
            bool occ = false;
#pragma unroll
            for (int k = 0; k < 8; ++k) occ |= r[k] >= 0;
            if (!occ) {
#pragma unroll
                for (int c = 0; c < 8; ++c) store(dst + c * plane, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f);
            } else {
                float v[8][8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
#pragma unroll
                    for (int c = 0; c < 8; ++c) v[k][c] = 0.f;
                    if (r[k] >= 0) ldg256_stream(feats + static_cast<int64_t>(r[k]) * f + c0, v[k]);
                }
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    store(dst + c * plane, v[0][c], v[1][c], v[2][c], v[3][c], v[4][c], v[5][c], v[6][c], v[7][c]);
            }
        }
        lo = nlo;
        hi = nhi;
        item += n_warps;
    }
}

// ---- variant 6 ----------------------------------------------------------------------------------
// Zero stream first, patch second.  Same work split as variant 4 (CTA = blockDim.x * 8 cells x 8 channels, plane-major), but a
// lane stores its 8 x 32 bytes of zeros as soon as its index-map entry has arrived and only then fetches the 32-byte
// feature sector of each occupied cell (5 % of the cells) and overwrites those elements with 4-byte stores, which merge
// into the lines the lane has just written while they are still dirty in L2.  The bulk write stream therefore never waits
// for the gather, and since no canvas data is held in registers the SM keeps three times as many warps in flight.
template <bool CS>
__global__ void __launch_bounds__(kThreads)
k_scatter_patch(const float *__restrict__ feats, const int32_t *__restrict__ cell_row, int f, int64_t plane,
                int tiles_per_plane, float *__restrict__ bev)
{
    const int groups = f >> 3;
    const int tile = blockIdx.x % tiles_per_plane;
    const int cg = (blockIdx.x / tiles_per_plane) % groups;
    const int b = blockIdx.x / (groups * tiles_per_plane);
    const int64_t cell0 = static_cast<int64_t>(tile) * (blockDim.x * 8) + threadIdx.x * 8;
    if (cell0 >= plane) return;
    const int4 *src = reinterpret_cast<const int4 *>(cell_row + b * plane + cell0);
    const int4 lo = __ldg(src), hi = __ldg(src + 1);
    const int c0 = cg * 8;
    float *dst = bev + (static_cast<int64_t>(b) * f + c0) * plane + cell0;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        if (CS) stg256_cs(dst + c * plane, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f);
        else stg256(dst + c * plane, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f);
    }
    const int32_t r[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        if (r[k] >= 0) {
            float v[8];
            ldg256_stream(feats + static_cast<int64_t>(r[k]) * f + c0, v);
#pragma unroll
            for (int c = 0; c < 8; ++c) dst[c * plane + k] = v[c];
        }
    }
}

// ---- fp16 canvas ---------------------------------------------------------------------------------
// The product's caller stores the BEV map as float16 (src/get-data/precompute_bev_features.py:394: bev.astype(np.float16));
// converting in the scatter halves the dominant write.  Same work split as variant 4: a lane owns 8 cells = one 16-byte
// store per channel, round-to-nearest-even like numpy / torch.
__global__ void __launch_bounds__(kThreads)
k_scatter_wide_half(const float *__restrict__ feats, const int32_t *__restrict__ cell_row, int f, int64_t plane,
                    int tiles_per_plane, __half *__restrict__ bev)
{
    const int groups = f >> 3;
    const int tile = blockIdx.x % tiles_per_plane;
    const int cg = (blockIdx.x / tiles_per_plane) % groups;
    const int b = blockIdx.x / (groups * tiles_per_plane);
    const int64_t cell0 = static_cast<int64_t>(tile) * (blockDim.x * 8) + threadIdx.x * 8;
    const bool inb = cell0 < plane;
    int32_t r[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) r[k] = -1;
    if (inb) {
        const int4 *src = reinterpret_cast<const int4 *>(cell_row + b * plane + cell0);
        const int4 v = __ldg(src), w = __ldg(src + 1);
        r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
        r[4] = w.x; r[5] = w.y; r[6] = w.z; r[7] = w.w;
    }
    bool any = false;
#pragma unroll
    for (int k = 0; k < 8; ++k) any |= r[k] >= 0;
    const int c0 = cg * 8;
    __half *dst = bev + (static_cast<int64_t>(b) * f + c0) * plane + cell0;
    if (!__any_sync(0xffffffffu, any)) {
        if (inb) {
#pragma unroll
            for (int c = 0; c < 8; ++c) __stcs(reinterpret_cast<uint4 *>(dst + c * plane), make_uint4(0u, 0u, 0u, 0u));
        }
        return;
    }
    if (!inb) return;
    float v[8][8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
#pragma unroll
        for (int c = 0; c < 8; ++c) v[k][c] = 0.f;
        if (r[k] >= 0) ldg256_stream(feats + static_cast<int64_t>(r[k]) * f + c0, v[k]);
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        uint4 o;
        __half2 h;
        h = __floats2half2_rn(v[0][c], v[1][c]); o.x = *reinterpret_cast<uint32_t *>(&h);
        h = __floats2half2_rn(v[2][c], v[3][c]); o.y = *reinterpret_cast<uint32_t *>(&h);
        h = __floats2half2_rn(v[4][c], v[5][c]); o.z = *reinterpret_cast<uint32_t *>(&h);
        h = __floats2half2_rn(v[6][c], v[7][c]); o.w = *reinterpret_cast<uint32_t *>(&h);
        *reinterpret_cast<uint4 *>(dst + c * plane) = o;
    }
}

// ---- async-proxy helpers ------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// The canvas is written once and not read by this path: mark its lines evict-first so the 1 GiB stream does not push the
// index map and the pillar features (both re-read by this very kernel) out of the 126 MB L2.
__device__ __forceinline__ uint64_t policy_evict_first()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_store_1d(void *gdst, const void *ssrc, uint32_t bytes, uint64_t pol)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gdst),
                 "r"(smem_u32(ssrc)), "r"(bytes), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *tmap, const void *ssrc, int c0, int c1, uint64_t pol)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;" ::"l"(
                     reinterpret_cast<uint64_t>(tmap)),
                 "r"(smem_u32(ssrc)), "r"(c0), "r"(c1), "l"(pol)
                 : "memory");
}

__device__ __forceinline__ void bulk_store_1d_nohint(void *gdst, const void *ssrc, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void tma_store_2d_nohint(const CUtensorMap *tmap, const void *ssrc, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(tmap)),
                 "r"(smem_u32(ssrc)), "r"(c0), "r"(c1)
                 : "memory");
}

// ---- variants 2 and 3 ---------------------------------------------------------------------------
// Persistent CTAs; two shared-memory tiles [f][kCells] per CTA, alternating.  Invariant: a tile that is not "dirty" is
// all zero, so an empty stretch of the canvas is stored straight from it with no shared-memory write and no wait at all.
// A tile that received pillar columns is restored lazily, the next time the buffer comes round (two tiles later), when the
// store that read it has long finished (cp.async.bulk.wait_group.read 1).  The index-map entry of the next tile is
// prefetched one iteration ahead so its latency hides behind the current tile.
__device__ __forceinline__ void bulk_wait_read_1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }

template <bool TMA2D, int kCells, bool kHint>
__global__ void __launch_bounds__(kThreads)
k_scatter_async(const float *__restrict__ feats, const int32_t *__restrict__ cell_row, int f, int64_t plane,
                int tiles_per_plane, int64_t n_tiles, float *__restrict__ bev, const __grid_constant__ CUtensorMap tmap)
{
    extern __shared__ __align__(128) float s_tiles[];  // [2][f][kCells]
    const int tid = threadIdx.x;
    const int tile_floats = f * kCells;
    for (int i = tid; i < 2 * tile_floats; i += kThreads) s_tiles[i] = 0.f;
    fence_proxy_async();
    __syncthreads();

    constexpr int kParts = kThreads / kCells;  // threads per cell
    const int cell = tid % kCells, part = tid / kCells;
    const int c_lo = part * (f / kParts), c_hi = c_lo + f / kParts;
    const bool issuer = TMA2D ? (tid == 0) : (tid < f);
    const uint64_t pol = kHint ? policy_evict_first() : 0ull;

    bool dirty[2] = {false, false};       // CTA-uniform
    int32_t r_prev[2] = {-1, -1};         // the row this thread copied into buffer k last time
    int64_t t = blockIdx.x;
    int32_t r_next = -1;
    if (t < n_tiles) {
        const int b = static_cast<int>(t / tiles_per_plane);
        const int64_t cell0 = (t % tiles_per_plane) * kCells;
        if (cell0 + cell < plane) r_next = __ldg(cell_row + b * plane + cell0 + cell);
    }
    for (int it = 0; t < n_tiles; t += gridDim.x, ++it) {
        const int k = it & 1;
        float *tile = s_tiles + k * tile_floats;
        const int b = static_cast<int>(t / tiles_per_plane);
        const int64_t cell0 = (t % tiles_per_plane) * kCells;
        const int ncell = static_cast<int>(tmin<int64_t>(kCells, plane - cell0));
        const int32_t r = r_next;
        r_next = -1;
        const int64_t tn = t + gridDim.x;
        if (tn < n_tiles) {
            const int bn = static_cast<int>(tn / tiles_per_plane);
            const int64_t celln = (tn % tiles_per_plane) * kCells;
            if (celln + cell < plane) r_next = __ldg(cell_row + bn * plane + celln + cell);
        }
        const int any = __syncthreads_or(r >= 0);
        if (any || dirty[k]) {
            if (issuer) bulk_wait_read_1();  // the store issued from this buffer two tiles ago has released it
            __syncthreads();
            if (dirty[k] && r_prev[k] >= 0)
                for (int c = c_lo; c < c_hi; ++c) tile[c * kCells + cell] = 0.f;
            if (r >= 0) {
                const float4 *src = reinterpret_cast<const float4 *>(feats + static_cast<int64_t>(r) * f + c_lo);
                for (int c = c_lo; c < c_hi; c += 4) {
                    const float4 v = __ldg(src++);
                    tile[(c + 0) * kCells + cell] = v.x;
                    tile[(c + 1) * kCells + cell] = v.y;
                    tile[(c + 2) * kCells + cell] = v.z;
                    tile[(c + 3) * kCells + cell] = v.w;
                }
            }
            fence_proxy_async();
            __syncthreads();
            dirty[k] = any != 0;
            r_prev[k] = r;
        }
        if (TMA2D) {
            if (tid == 0) {
                if (kHint) tma_store_2d(&tmap, tile, static_cast<int>(cell0), b * f, pol);
                else tma_store_2d_nohint(&tmap, tile, static_cast<int>(cell0), b * f);
                bulk_commit();
            }
        } else if (tid < f) {
            if (kHint)
                bulk_store_1d(bev + (static_cast<int64_t>(b) * f + tid) * plane + cell0, tile + tid * kCells,
                              static_cast<uint32_t>(ncell) * 4u, pol);
            else
                bulk_store_1d_nohint(bev + (static_cast<int64_t>(b) * f + tid) * plane + cell0, tile + tid * kCells,
                                     static_cast<uint32_t>(ncell) * 4u);
            bulk_commit();
        }
    }
    if (issuer) bulk_wait_all();
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

int sm_count()
{
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    return sms;
}

}  // namespace

cudaError_t launch_build_cell_row(const void *coords, bool coords_float, int64_t m, const int32_t *m_dev, int nb, int nx,
                                  int ny, int nz, int32_t *cell_row, cudaStream_t st)
{
    cudaError_t err = cudaMemsetAsync(cell_row, 0xFF, sizeof(int32_t) * static_cast<size_t>(nb) * nx * ny * nz, st);
    note_launch();
    if (err != cudaSuccess || m == 0) return err;
    const unsigned blocks = static_cast<unsigned>((m + kThreads - 1) / kThreads);
    k_build_cell_row<<<blocks, kThreads, 0, st>>>(coords, coords_float ? 1 : 0, m, m_dev, nb, nx, ny, nz, cell_row);
    note_launch();
    return cudaGetLastError();
}

cudaError_t launch_scatter_half(const float *feats, const int32_t *cell_row, int nb, int f, int nx, int ny, void *bev,
                                cudaStream_t st)
{
    const int64_t plane = static_cast<int64_t>(nx) * ny;
    if (nb == 0 || plane == 0 || f == 0) return cudaSuccess;
    if (plane % 8 != 0 || f % 8 != 0 || reinterpret_cast<uintptr_t>(bev) % 16 != 0 ||
        reinterpret_cast<uintptr_t>(feats) % 32 != 0)
        return cudaErrorInvalidValue;
    const int bs = 128;
    const int tpp = static_cast<int>((plane + bs * 8 - 1) / (bs * 8));
    const unsigned grid = static_cast<unsigned>(nb) * tpp * (f / 8);
    k_scatter_wide_half<<<grid, bs, 0, st>>>(feats, cell_row, f, plane, tpp, static_cast<__half *>(bev));
    note_launch();
    return cudaGetLastError();
}

cudaError_t launch_scatter(const float *feats, const int32_t *cell_row, int nb, int f, int nx, int ny, float *bev,
                           int variant, cudaStream_t st)
{
    const int64_t plane = static_cast<int64_t>(nx) * ny;
    if (nb == 0 || plane == 0 || f == 0) return cudaSuccess;
    const bool vec_ok = (plane % 4 == 0) && (reinterpret_cast<uintptr_t>(bev) % 16 == 0) &&
                        (reinterpret_cast<uintptr_t>(feats) % 16 == 0) && (f % 8 == 0);
    const size_t smem = sizeof(float) * 2 * f * 128;
    const bool async_ok = vec_ok && f <= kThreads && f % 8 == 0 && smem <= 200 * 1024;
    const bool wide_ok = vec_ok && (plane % 8 == 0) && (reinterpret_cast<uintptr_t>(bev) % 32 == 0);
    // measured on B200, cfg2 (profiles/r01_scatter_variants.md): channel-group 256-bit stores 172 us, direct stores with all
    // channels per warp 203 us, TMA tile stores 268 us
    if (variant == 0) variant = wide_ok ? 4 : 1;
    if ((variant == 2 || variant == 3) && !async_ok) variant = 1;
    if (variant == 3 && !get_encode_fn()) variant = 2;
    if (variant == 4 && !wide_ok) variant = 1;

    if (variant == 5 && !(wide_ok && f % 8 == 0)) variant = 1;
    if (variant == 5) {
        static int cs = -1, bs = 0, per_sm = 0;
        if (cs < 0) {
            const char *e1 = getenv("PILLARS_SCATTER_PERSIST_CS"), *e2 = getenv("PILLARS_SCATTER_PERSIST_BLOCK"),
                       *e3 = getenv("PILLARS_SCATTER_PERSIST_CTAS");
            cs = e1 ? atoi(e1) : 0;
            bs = e2 ? atoi(e2) : 128;
            per_sm = e3 ? atoi(e3) : 4;
            if (bs != 32 && bs != 64 && bs != 128 && bs != 256) bs = 128;
            if (per_sm < 1 || per_sm > 32) per_sm = 4;
        }
        const int tpp = static_cast<int>((plane + 255) / 256);
        const int64_t n_items = static_cast<int64_t>(nb) * (f / 8) * tpp;
        const int64_t ctas_needed = (n_items * 32 + bs - 1) / bs;
        const unsigned grid = static_cast<unsigned>(tmin<int64_t>(ctas_needed, static_cast<int64_t>(sm_count()) * per_sm));
        if (cs) k_scatter_persist<true><<<grid, bs, 0, st>>>(feats, cell_row, f, plane, tpp, n_items, bev);
        else k_scatter_persist<false><<<grid, bs, 0, st>>>(feats, cell_row, f, plane, tpp, n_items, bev);
        note_launch();
        return cudaGetLastError();
    }
    if (variant == 6 && !(wide_ok && f % 8 == 0)) variant = 1;
    if (variant == 6) {
        static int cs = -1, bs = 0;
        if (cs < 0) {
            const char *e1 = getenv("PILLARS_SCATTER_PATCH_CS"), *e2 = getenv("PILLARS_SCATTER_PATCH_BLOCK");
            cs = e1 ? atoi(e1) : 0;
            bs = e2 ? atoi(e2) : 128;
            if (bs != 32 && bs != 64 && bs != 128 && bs != 256) bs = 128;
        }
        const int tpp = static_cast<int>((plane + bs * 8 - 1) / (bs * 8));
        const unsigned grid = static_cast<unsigned>(nb) * tpp * (f / 8);
        if (cs) k_scatter_patch<true><<<grid, bs, 0, st>>>(feats, cell_row, f, plane, tpp, bev);
        else k_scatter_patch<false><<<grid, bs, 0, st>>>(feats, cell_row, f, plane, tpp, bev);
        note_launch();
        return cudaGetLastError();
    }
    if (variant == 4) {
        static int cs = -1, bs = 0, cpc = 0, tile_major = 0;
        if (cs < 0) {
            const char *e1 = getenv("PILLARS_SCATTER_WIDE_CS"), *e2 = getenv("PILLARS_SCATTER_WIDE_BLOCK"),
                       *e3 = getenv("PILLARS_SCATTER_WIDE_CHAN"), *e4 = getenv("PILLARS_SCATTER_WIDE_TILEMAJOR");
            cs = e1 ? atoi(e1) : 0;
            bs = e2 ? atoi(e2) : 128;
            cpc = e3 ? atoi(e3) : 8;
            tile_major = e4 ? atoi(e4) : 0;
            if (bs != 32 && bs != 64 && bs != 128 && bs != 256) bs = 128;
        }
        int chan = (cpc >= 8 && cpc % 8 == 0 && f % cpc == 0) ? cpc : f;
        const int tpp = static_cast<int>((plane + bs * 8 - 1) / (bs * 8));
        const unsigned grid = static_cast<unsigned>(nb) * tpp * (f / chan);
        static int cpp = 0;
        if (!cpp) {
            const char *e5 = getenv("PILLARS_SCATTER_WIDE_CPP");
            cpp = (e5 && atoi(e5) == 4) ? 4 : 8;
        }
        if (cs && cpp == 8) k_scatter_wide<true, 8><<<grid, bs, 0, st>>>(feats, cell_row, f, plane, tpp, chan, tile_major, bev);
        else if (cs) k_scatter_wide<true, 4><<<grid, bs, 0, st>>>(feats, cell_row, f, plane, tpp, chan, tile_major, bev);
        else if (cpp == 8) k_scatter_wide<false, 8><<<grid, bs, 0, st>>>(feats, cell_row, f, plane, tpp, chan, tile_major, bev);
        else k_scatter_wide<false, 4><<<grid, bs, 0, st>>>(feats, cell_row, f, plane, tpp, chan, tile_major, bev);
        note_launch();
        return cudaGetLastError();
    }

    if (variant == 1) {
        static int mode = -1, bs = 0;
        if (mode < 0) {
            const char *e1 = getenv("PILLARS_SCATTER_PLAIN_MODE"), *e2 = getenv("PILLARS_SCATTER_PLAIN_BLOCK");
            mode = e1 ? atoi(e1) : 0;
            bs = e2 ? atoi(e2) : kThreads;
            if (bs != 64 && bs != 128 && bs != 256) bs = kThreads;
        }
        if (vec_ok) {
            const int tpp = static_cast<int>((plane + bs * 4 - 1) / (bs * 4));
            const unsigned grid = static_cast<unsigned>(nb) * tpp;
            if (mode == 1) k_scatter_plain<true, 4, false><<<grid, bs, 0, st>>>(feats, cell_row, f, plane, tpp, bev);
            else if (mode == 2) k_scatter_plain<true, 8, true><<<grid, bs, 0, st>>>(feats, cell_row, f, plane, tpp, bev);
            else if (mode == 3) k_scatter_plain<true, 4, true><<<grid, bs, 0, st>>>(feats, cell_row, f, plane, tpp, bev);
            else k_scatter_plain<true, 8, false><<<grid, bs, 0, st>>>(feats, cell_row, f, plane, tpp, bev);
        } else {
            const int tpp = static_cast<int>((plane + kThreads - 1) / kThreads);
            k_scatter_plain<false, 8, false><<<static_cast<unsigned>(nb) * tpp, kThreads, 0, st>>>(feats, cell_row, f, plane, tpp, bev);
        }
        note_launch();
        return cudaGetLastError();
    }

    // tuning knobs (measurement only): tile width, CTAs per SM, L2 hint
    static int env_cells = -1, env_ctas = -1, env_nohint = -1;
    if (env_cells < 0) {
        const char *e1 = getenv("PILLARS_SCATTER_CELLS"), *e2 = getenv("PILLARS_SCATTER_CTAS"), *e3 = getenv("PILLARS_SCATTER_NOHINT");
        env_cells = e1 ? atoi(e1) : 0;
        env_ctas = e2 ? atoi(e2) : 0;
        env_nohint = e3 ? atoi(e3) : 0;
    }
    const int cells = (env_cells == 256 || env_cells == 64) ? env_cells : 128;
    const size_t smem_bytes = sizeof(float) * 2 * f * cells;
    const int tpp = static_cast<int>((plane + cells - 1) / cells);
    const int64_t n_tiles = static_cast<int64_t>(nb) * tpp;
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    if (variant == 3) {
        const cuuint64_t gdim[2] = {static_cast<cuuint64_t>(plane), static_cast<cuuint64_t>(nb) * f};
        const cuuint64_t gstride[1] = {static_cast<cuuint64_t>(plane) * sizeof(float)};
        const cuuint32_t box[2] = {static_cast<cuuint32_t>(cells), static_cast<cuuint32_t>(f)};
        const cuuint32_t estr[2] = {1, 1};
        const CUresult r = get_encode_fn()(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, bev, gdim, gstride, box, estr,
                                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                           CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) variant = 2;
    }
    int per_sm = static_cast<int>((220 * 1024) / (smem_bytes + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 8) per_sm = 8;
    if (env_ctas > 0 && env_ctas < per_sm) per_sm = env_ctas;
    const unsigned grid = static_cast<unsigned>(tmin<int64_t>(n_tiles, static_cast<int64_t>(sm_count()) * per_sm));
#define PILLARS_LAUNCH_SCATTER(T2D, CELLS, HINT)                                                                          \
    do {                                                                                                                  \
        cudaFuncSetAttribute(k_scatter_async<T2D, CELLS, HINT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);  \
        k_scatter_async<T2D, CELLS, HINT><<<grid, kThreads, smem_bytes, st>>>(feats, cell_row, f, plane, tpp, n_tiles, bev, \
                                                                              tmap);                                      \
    } while (0)
#define PILLARS_LAUNCH_SCATTER_C(T2D, HINT)                      \
    do {                                                         \
        if (cells == 64) PILLARS_LAUNCH_SCATTER(T2D, 64, HINT);  \
        else if (cells == 256) PILLARS_LAUNCH_SCATTER(T2D, 256, HINT); \
        else PILLARS_LAUNCH_SCATTER(T2D, 128, HINT);             \
    } while (0)
    if (variant == 3) {
        if (env_nohint) PILLARS_LAUNCH_SCATTER_C(true, false);
        else PILLARS_LAUNCH_SCATTER_C(true, true);
    } else {
        if (env_nohint) PILLARS_LAUNCH_SCATTER_C(false, false);
        else PILLARS_LAUNCH_SCATTER_C(false, true);
    }
    note_launch();
    return cudaGetLastError();
}

}  // namespace pillars
