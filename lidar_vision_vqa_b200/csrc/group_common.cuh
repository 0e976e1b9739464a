// Pieces shared by the two point -> pillar grouping implementations (voxelize.cu: open-addressing hash table,
// group_dense.cu: direct-mapped cell table).
#pragma once

#include "common.cuh"

namespace pillars {

constexpr int kGroupThreads = 256;  // points per CTA of the insert / place kernels
constexpr int kLookGroup = 256;     // scan tiles per look-back group

__device__ __forceinline__ unsigned lanemask_lt()
{
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u32(uint32_t *p, uint32_t v)
{
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ uint32_t div_by(uint32_t v, uint32_t d, int sh) { return sh >= 0 ? v >> sh : v / d; }

inline int log2_exact(uint32_t v)
{
    if (v == 0 || (v & (v - 1)) != 0) return -1;
    int s = 0;
    while ((1u << s) != v) ++s;
    return s;
}

// One contiguous, fully coalesced read of a tile of `count` rows into shared memory (rows are 12..64 B, so per-row vector
// loads would not be coalesced).
__device__ __forceinline__ void load_point_tile(const float *__restrict__ src, int count, int stride, int vec_ok,
                                                float *__restrict__ s_pts, int tid, int threads)
{
    const int nfl = count * stride;
    if (vec_ok) {
        const float4 *src4 = reinterpret_cast<const float4 *>(src);
        float4 *dst4 = reinterpret_cast<float4 *>(s_pts);
        const int n4 = nfl >> 2;
        for (int i = tid; i < n4; i += threads) dst4[i] = __ldg(src4 + i);
        for (int i = (n4 << 2) + tid; i < nfl; i += threads) s_pts[i] = __ldg(src + i);
    } else {
        for (int i = tid; i < nfl; i += threads) s_pts[i] = __ldg(src + i);
    }
}

// Frames touched by a tile of points [lo_i, hi_i]: this thread's share of (number of frame starts <= lo_i / hi_i); summed
// over the CTA, the frame index is that count minus one.  Every load is independent (a binary search would chain
// ~log2(nb) dependent global round trips in front of the CTA's first barrier).
__device__ __forceinline__ void count_frame_starts(const int32_t *__restrict__ frame_offsets, int nb, int64_t lo_i, int64_t hi_i,
                                                   int tid, int threads, uint32_t &c_lo, uint32_t &c_hi)
{
    c_lo = 0;
    c_hi = 0;
    for (int f = tid; f < nb; f += threads) {
        const int64_t o = __ldg(frame_offsets + f);
        c_lo += o <= lo_i ? 1u : 0u;
        c_hi += o <= hi_i ? 1u : 0u;
    }
}

// frame holding point i: largest b with offsets[b] <= i (offsets[0] = 0, offsets[nb] = n)
__device__ __forceinline__ int find_frame(const int32_t *__restrict__ frame_offsets, int nb, int64_t i)
{
    int lo = 0, hi = nb;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(frame_offsets + mid) <= i) lo = mid; else hi = mid;
    }
    return lo;
}

// IEEE fp32 sub, true division, floor, no contraction: bit-identical to the CPU voxeliser.  Returns false when the point
// falls outside the grid (NaN / Inf coordinates fail the comparisons).
__device__ __forceinline__ bool quantize_point(const float *__restrict__ p, const GridDev &gd, uint32_t &cell)
{
    const float fx = floorf(__fdiv_rn(__fsub_rn(p[0], gd.rmin[0]), gd.vsz[0]));
    const float fy = floorf(__fdiv_rn(__fsub_rn(p[1], gd.rmin[1]), gd.vsz[1]));
    const float fz = gd.ignore_z ? 0.f : floorf(__fdiv_rn(__fsub_rn(p[2], gd.rmin[2]), gd.vsz[2]));
    const bool valid = (fx >= 0.f) && (fx < static_cast<float>(gd.g[0])) && (fy >= 0.f) && (fy < static_cast<float>(gd.g[1])) &&
                       (fz >= 0.f) && (fz < static_cast<float>(gd.g[2]));
    if (valid) {
        const uint32_t cx = static_cast<uint32_t>(fx), cy = static_cast<uint32_t>(fy), cz = static_cast<uint32_t>(fz);
        cell = (cz * gd.g[1] + cy) * gd.g[0] + cx;
    }
    return valid;
}

// What the place stage needs to turn a point into its record
struct PlaceParams {
    const float *points;
    int64_t n;
    int stride, col0, c_point;
    const int32_t *point_slot;       // hash: slot of the point; dense: its cell key (b * cells + cell); -1 = rejected
    const uint32_t *point_arrival;
    const HashEntry *table;          // hash path
    const uint32_t *cell_first;      // dense path: tagged list base per cell
    uint32_t *sorted_idx;            // or NULL
    PointRecord *records;            // or NULL
    uint4 *pillar_meta;
    uint4 *long_list;
    uint32_t *long_count;
    const uint32_t *pillar_cnt, *frame_gstart, *frame_rowbase;
    GridDev gd;
    int sh_cells, sh_cells_xy, sh_nx;  // log2 of the divisor when it is a power of two, else -1
    float vsz[3], off[3];
    int32_t *voxel_coords, *voxel_num_points, *cell_row;
    int64_t capacity;
    unsigned long long *dbg;
};

inline PlaceParams make_place_params(const float *points, int64_t n, int stride, int col0, int c_point, const GridDev &gd,
                                     const Workspace &ws, bool want_index_lists, const PlaceExtras &extras)
{
    PlaceParams pp{};
    pp.points = points;
    pp.n = n;
    pp.stride = stride;
    pp.col0 = col0;
    pp.c_point = c_point;
    pp.point_slot = ws.point_slot;
    pp.point_arrival = ws.point_arrival;
    pp.table = ws.table;
    pp.cell_first = ws.cell_first;
    pp.sorted_idx = want_index_lists ? ws.sorted_idx : nullptr;
    pp.records = extras.records ? ws.records : nullptr;
    pp.pillar_meta = ws.pillar_meta;
    pp.long_list = ws.long_list;
    pp.long_count = ws.long_count;
    pp.pillar_cnt = ws.pillar_cnt;
    pp.frame_gstart = ws.frame_gstart;
    pp.frame_rowbase = ws.frame_rowbase;
    pp.gd = gd;
    pp.sh_cells = log2_exact(gd.cells);
    pp.sh_cells_xy = log2_exact(gd.cells_xy);
    pp.sh_nx = log2_exact(static_cast<uint32_t>(gd.g[0]));
    for (int k = 0; k < 3; ++k) {
        pp.vsz[k] = extras.vsz[k];
        pp.off[k] = extras.off[k];
    }
    pp.voxel_coords = extras.records ? extras.voxel_coords : nullptr;
    pp.voxel_num_points = extras.records ? extras.voxel_num_points : nullptr;
    pp.cell_row = extras.records && extras.write_cell_row ? ws.cell_row : nullptr;
    pp.capacity = extras.capacity;
    pp.dbg = debug_times_ptr();
    return pp;
}

struct CellCoord {
    uint32_t b, z, y, x;
};

// The per-pillar record of the streaming feature kernel, at the pillar's list start position; pillars that cannot fit its
// 64-position window (more than 32 points) are also appended to a list that its warps drain after their own chunks.
__device__ __forceinline__ void publish_pillar(uint4 *pillar_meta, uint4 *long_list, uint32_t *long_count, uint32_t base,
                                               const CellCoord &c, uint32_t row_or_ff, uint32_t n)
{
    const uint32_t xy = c.x | (c.y << 16);
    pillar_meta[base] = make_uint4(xy, row_or_ff, n, c.z);
    if (n > 32u) {
        const uint32_t slot = atomicAdd(long_count, 1u) + 1u;  // the counter starts at 0xFFFFFFFF
        long_list[2 * slot] = make_uint4(base, n, row_or_ff, xy);
        long_list[2 * slot + 1] = make_uint4(c.z, 0u, 0u, 0u);
    }
}

template <typename P>
__device__ __forceinline__ CellCoord decode_key(const P &p, uint32_t key)
{
    CellCoord c;
    c.b = div_by(key, p.gd.cells, p.sh_cells);
    const uint32_t cell = key - c.b * p.gd.cells;
    c.z = div_by(cell, p.gd.cells_xy, p.sh_cells_xy);
    const uint32_t rem = cell - c.z * p.gd.cells_xy;
    c.y = div_by(rem, static_cast<uint32_t>(p.gd.g[0]), p.sh_nx);
    c.x = rem - c.y * static_cast<uint32_t>(p.gd.g[0]);
    return c;
}

// The point's 32-byte record: coordinates relative to the pillar centre (coord * voxel + offset, two roundings as in the
// reference, pillar_vfe.py:100-103), intensity, time, its index and its position inside the pillar's list.
template <typename P>
__device__ __forceinline__ void make_record(const P &p, int64_t i, const CellCoord &c, uint32_t arrival, float4 &a, float4 &d)
{
    const float cx = __fadd_rn(__fmul_rn(static_cast<float>(c.x), p.vsz[0]), p.off[0]);
    const float cy = __fadd_rn(__fmul_rn(static_cast<float>(c.y), p.vsz[1]), p.off[1]);
    const float cz = __fadd_rn(__fmul_rn(static_cast<float>(c.z), p.vsz[2]), p.off[2]);
    const float *q = p.points + i * p.stride + p.col0;
    a.x = __fsub_rn(__ldg(q), cx);
    a.y = __fsub_rn(__ldg(q + 1), cy);
    a.z = __fsub_rn(__ldg(q + 2), cz);
    a.w = p.c_point > 3 ? __ldg(q + 3) : 0.f;
    d.x = p.c_point > 4 ? __ldg(q + 4) : 0.f;
    d.y = 0.f;  // walk-control flags, set by the feature kernel in its staged copy
    d.z = __uint_as_float(static_cast<uint32_t>(i));
    d.w = __uint_as_float(arrival);
}

template <typename P>
__device__ __forceinline__ void write_record(const P &p, int64_t i, const CellCoord &c, uint32_t pos, uint32_t arrival)
{
    float4 a, d;
    make_record(p, i, c, arrival, a, d);
    float4 *dst = reinterpret_cast<float4 *>(p.records + pos);
    dst[0] = a;
    dst[1] = d;
}

}  // namespace pillars
