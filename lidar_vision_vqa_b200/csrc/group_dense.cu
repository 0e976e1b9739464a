// Point -> pillar grouping with a DIRECT-MAPPED cell table (same semantics and same workspace products as voxelize.cu;
// replaces the spconv CPU call behind src/lidar-encoder/pcdet/datasets/processor/data_processor.py:16-61,133-180).
//
// Every pillar grid of the reference's configs has n_frames * cells within a small multiple of the point count (16 x 512^2
// cells for 0.5 M points), so the "hash table" can be the identity: one 8-byte entry per (frame, cell), split into two
// 4-byte planes that a single 0xFF memset initialises and that stay L2 resident.
//
//   k_insert_dense   coalesced tile loads -> fp32 sub/div/floor -> key = frame * cells + cell.  Adjacent lanes that hit the
//                    same cell (unshuffled sweeps) are merged; the run's first lane issues ONE atomicAdd on the count plane
//                    (returns the arrival rank of the run) and one fire-and-forget atomicMin on the first-index plane.  No
//                    probing, no CAS loop, no key compare: one dependent L2 round trip per point instead of two or more.
//                    The claim of an empty cell also counts the frame's pillars (one atomic per CTA), so the scan kernel
//                    knows every frame's row base before it starts.
//   k_scan_dense     single-pass chained scan over points in index order of [point is the first of its cell] and of the
//                    cell counts => pillar id in first-appearance order + start of its point list (no sort).  The thread
//                    that owns a pillar's first point writes everything per-pillar that needs no feature arithmetic:
//                    voxel_coords, voxel_num_points, the pillar's {x, y, z, row, n} record for the feature kernel, the BEV
//                    index-map entry (the count plane BECOMES the index map: empty cells already hold -1), and the
//                    list base into the first-index plane (tagged with bit 31 so that it can never be mistaken for a
//                    point index by a tile that is still testing "am I the first").
//   k_place_dense    point -> its pillar's list (index list and/or 32-byte records): key -> base -> position.
#include "group_common.cuh"


namespace pillars {

namespace {

constexpr int kScanThreads = 512;
constexpr int kScanPer = kTile / kScanThreads;  // 2
constexpr uint32_t kDenseLookGroup = kScanThreads;  // scan tiles per look-back group: one descriptor per thread

constexpr uint32_t kBaseTag = 0x80000000u;
constexpr unsigned kFullMask = 0xffffffffu;

// 0xFF fill of the cell table (+ counters and scan descriptors in front of it).  A kernel rather than a memset so that the
// insert kernel can be launched behind it programmatically: its CTAs load and quantise their points while the fill drains.
__global__ void __launch_bounds__(256) k_fill_ff(uint4 *__restrict__ dst, size_t n_vec, unsigned long long *dbg)
{
    pdl_trigger();
    if (threadIdx.x == 0) dbg_stamp(dbg, 0);
    const uint4 ff = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_vec; i += stride) dst[i] = ff;
    if (threadIdx.x == 0) dbg_stamp(dbg, 1);
}

constexpr int kInsPer = 2;                          // points per thread: two independent atomic round trips in flight
constexpr int kInsTile = kGroupThreads * kInsPer;   // points per CTA

__global__ void __launch_bounds__(kGroupThreads, 8)
k_insert_dense(const float *__restrict__ points, int64_t n, int stride, int col0, const int32_t *__restrict__ frame_offsets,
               int nb, GridDev gd, uint32_t *__restrict__ cell_first, uint32_t *__restrict__ cell_cnt,
               int32_t *__restrict__ point_key, uint32_t *__restrict__ point_arrival, uint32_t *__restrict__ frame_new,
               int vec_ok, Header *hdr, unsigned long long *dbg)
{
    extern __shared__ __align__(16) float s_pts[];  // [kInsTile * stride]
    __shared__ uint32_t s_cnt[2];
    __shared__ uint32_t s_claims;

    const int tid = threadIdx.x, lane = tid & 31;
    // The scan kernel's CTAs draw their tile ids from a counter that is never reset: CTA 0 notes where it stands before any
    // of them can exist (they are launched once every CTA of this grid has passed the trigger below).
    if (blockIdx.x == 0) {
        if (tid == 0) {
            st_relaxed_u32(&hdr->ticket_base, ld_relaxed_u32(&hdr->ticket_seq));
            __threadfence();
        }
        __syncthreads();  // no thread of this CTA triggers before the base is out
    }
    pdl_trigger();  // the scan kernel may start taking SMs as this grid's CTAs retire
    const int64_t tile_start = static_cast<int64_t>(blockIdx.x) * kInsTile;
    const int count = static_cast<int>(tmin<int64_t>(kInsTile, n - tile_start));
    if (tid < 2) s_cnt[tid] = 0u;
    if (tid == 2) s_claims = 0u;
    if (tid == 0) {
        dbg_stamp(dbg, 2);
        dbg_stamp(dbg, 25);  // latest CTA start
    }
    __syncthreads();
    // (the points and the frame offsets are inputs of the call: no dependency on the fill kernel yet)
    load_point_tile(points + tile_start * stride, count, stride, vec_ok, s_pts, tid, kGroupThreads);
    {  // frames touched by this tile: [b0, b1] = (frame starts <= first / last point of the tile) - 1; independent loads
        uint32_t c_lo, c_hi;
        count_frame_starts(frame_offsets, nb, tile_start, tile_start + count - 1, tid, kGroupThreads, c_lo, c_hi);
        if (c_lo) atomicAdd(&s_cnt[0], c_lo);
        if (c_hi) atomicAdd(&s_cnt[1], c_hi);
    }
    __syncthreads();
    if (tid == 0) dbg_stamp(dbg, 27);  // latest tile loaded
    const int s_b0 = static_cast<int>(s_cnt[0]) - 1, s_b1 = static_cast<int>(s_cnt[1]) - 1;

    bool valid[kInsPer];
    uint32_t key[kInsPer];
    int fb[kInsPer];
#pragma unroll
    for (int k = 0; k < kInsPer; ++k) {
        const int t = tid + k * kGroupThreads;  // consecutive lanes hold consecutive points in every round
        valid[k] = false;
        key[k] = 0;
        fb[k] = s_b0;
        if (t < count) {
            uint32_t cell = 0;
            valid[k] = quantize_point(s_pts + t * stride + col0, gd, cell);
            if (valid[k]) {
                const int64_t i = tile_start + t;
                while (fb[k] < s_b1 && __ldg(frame_offsets + fb[k] + 1) <= i) ++fb[k];
                key[k] = static_cast<uint32_t>(fb[k]) * gd.cells + cell;
            }
        }
    }
    if (tid == 0) dbg_stamp(dbg, 3);  // tile loaded and quantised
    pdl_wait();  // the table must be filled before the first atomic
    if (tid == 0) dbg_stamp(dbg, 4);  // (earliest) fill complete
    // Lanes hold consecutive points, so an unshuffled sweep puts the points of a cell on ADJACENT lanes: runs of equal keys
    // are merged (one atomic per run; two shuffles and a ballot -- __match_any_sync would also merge non-adjacent
    // duplicates, but costs ~3 us per call on the MIO path and a shuffled sweep has next to none).
    int head_lane[kInsPer];
    uint32_t old[kInsPer];
#pragma unroll
    for (int k = 0; k < kInsPer; ++k) {  // both rounds' atomics are issued before either result is used
        const uint32_t kk = valid[k] ? key[k] : 0xFFFFFFFFu - static_cast<uint32_t>(lane);  // invalid lanes: unique keys
        const uint32_t prev = __shfl_up_sync(kFullMask, kk, 1);
        const unsigned heads = __ballot_sync(kFullMask, lane == 0 || kk != prev);
        head_lane[k] = 31 - __clz(heads & (lanemask_lt() | (1u << lane)));
        const unsigned above = head_lane[k] == 31 ? 0u : heads & ~((2u << head_lane[k]) - 1u);
        const int run = (above ? __ffs(above) - 1 : 32) - head_lane[k];
        old[k] = 0;
        if (valid[k] && lane == head_lane[k]) {  // lowest lane == smallest point index of the run
            // count plane holds "points - 1" (0xFFFFFFFF = empty): the returned value + 1 is the run's arrival rank
            old[k] = atomicAdd(&cell_cnt[key[k]], static_cast<uint32_t>(run));
            atomicMin(&cell_first[key[k]], static_cast<uint32_t>(tile_start + tid + k * kGroupThreads));
        }
    }
    const bool one_frame = s_b0 == s_b1;
#pragma unroll
    for (int k = 0; k < kInsPer; ++k) {
        const int t = tid + k * kGroupThreads;
        const uint32_t head_old = __shfl_sync(kFullMask, old[k], head_lane[k]);
        const bool claimed = valid[k] && lane == head_lane[k] && old[k] == 0xFFFFFFFFu;
        // pillars opened per frame: one atomic per CTA when the tile lies inside one frame (the rule), else one per claim
        const unsigned cl = __ballot_sync(kFullMask, claimed);
        if (one_frame) {
            if (lane == 0 && cl) atomicAdd(&s_claims, static_cast<uint32_t>(__popc(cl)));
        } else if (claimed) {
            atomicAdd(&frame_new[fb[k]], 1u);
        }
        if (t < count) {
            point_key[tile_start + t] = valid[k] ? static_cast<int32_t>(key[k]) : -1;
            point_arrival[tile_start + t] = head_old + 1u + static_cast<uint32_t>(lane - head_lane[k]);
        }
    }
    __syncthreads();
    if (one_frame && tid == 0 && s_claims) atomicAdd(&frame_new[s_b0], s_claims);
    if (tid == 0) dbg_stamp(dbg, 5);
}

// ---------------------------------------------------------------------------------------------
// scan.  Descriptor = pillars(31) << 31 | listed points(31), bit 63 CLEAR when ready (the region is 0xFF-initialised);
// running sums travel as (pillars << 32 | listed)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long pack_desc(unsigned long long v)
{
    return ((v >> 32) << 31) | (v & 0x7FFFFFFFull);
}
__device__ __forceinline__ unsigned long long wait_desc(const unsigned long long *p)
{
    unsigned long long d = ld_relaxed_u64(p);
    while (d >> 63) d = ld_relaxed_u64(p);  // (a __nanosleep back-off between polls measured no better: 0 / 100 / 300 ns)
    return ((d >> 31) << 32) | (d & 0x7FFFFFFFull);
}

template <int kWarps>
__device__ __forceinline__ unsigned long long block_scan_excl(unsigned long long v, unsigned long long *s_warp, int lane,
                                                              int warp, unsigned long long &total)
{
    unsigned long long incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long o = __shfl_up_sync(kFullMask, incl, d);
        if (lane >= d) incl += o;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    unsigned long long wex = 0;
    total = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
        const unsigned long long s = s_warp[w];
        if (w < warp) wex += s;
        total += s;
    }
    __syncthreads();
    return wex + incl - v;
}

struct ScanDenseParams {
    int64_t n;
    uint32_t n_tiles;
    const int32_t *point_key;
    uint32_t *cell_first;
    int32_t *cell_row;  // count plane on entry, BEV index map on exit
    unsigned long long *tile_agg, *tile_prefix;
    const uint32_t *frame_new;
    int nb;
    Header *hdr;
    uint32_t *frame_gstart, *frame_rowbase;
    int32_t *pillar_count;
    uint32_t *pillar_key, *pillar_list, *pillar_cnt;  // when write_lists
    uint4 *pillar_meta, *long_list;                   // when write_meta
    uint32_t *long_count;
    int32_t *voxel_coords, *voxel_num_points;         // when write_meta (may be NULL)
    GridDev gd;
    int sh_cells, sh_cells_xy, sh_nx;
    int64_t capacity;
    int write_lists, write_meta;
    unsigned long long *dbg;
};

__global__ void __launch_bounds__(kScanThreads, 4) k_scan_dense(const __grid_constant__ ScanDenseParams p)
{
    constexpr int kWarps = kScanThreads / 32;
    __shared__ uint32_t s_tile;
    __shared__ unsigned long long s_warp[kWarps];
    __shared__ unsigned long long s_look[kWarps];
    __shared__ uint32_t s_gstart[kMaxFrames + 1], s_rowbase[kMaxFrames + 1];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) dbg_stamp(p.dbg, 6);   // CTA resident
    // A tile spins on its predecessors, so they must be running: ids are handed out in scheduling order.  The ticket is
    // drawn while the insert kernel is still running (the CTAs arrive spread out in time; hundreds of them hitting one
    // address right after the wait took 4 us to serve), from a sequence that only ever counts up.
    if (tid == 0) s_tile = atomicAdd(&p.hdr->ticket_seq, 1u) - ld_relaxed_u32(&p.hdr->ticket_base);
    pdl_wait();     // everything below reads what the insert kernel (and the fill before it) wrote
    pdl_trigger();
    if (tid == 0) dbg_stamp(p.dbg, 7);   // (latest) insert kernel complete
    if (tid == 0) dbg_stamp(p.dbg, 8);   // (earliest) insert kernel complete

    // per-frame pillar counts of the insert kernel: loaded now, scanned after this tile's aggregate is published
    const uint32_t mv = static_cast<uint32_t>(p.gd.max_voxels);
    const int f0 = 2 * tid;
    uint32_t g0 = 0, g1 = 0;
    if (f0 < p.nb) g0 = __ldcg(p.frame_new + f0) + 1u;
    if (f0 + 1 < p.nb) g1 = __ldcg(p.frame_new + f0 + 1) + 1u;
    __syncthreads();  // s_tile
    const uint32_t tile = s_tile;
    const int64_t tile_start = static_cast<int64_t>(tile) * kTile;
    const int64_t i0 = tile_start + tid * kScanPer;

    int32_t key[kScanPer];
    if (i0 + kScanPer <= p.n) {
        const int2 v = *reinterpret_cast<const int2 *>(p.point_key + i0);
        key[0] = v.x;
        key[1] = v.y;
    } else {
#pragma unroll
        for (int k = 0; k < kScanPer; ++k) key[k] = (i0 + k < p.n) ? p.point_key[i0 + k] : -1;
    }
    uint32_t first[kScanPer], cntm1[kScanPer];
#pragma unroll
    for (int k = 0; k < kScanPer; ++k)  // all gathers in flight before the first use
        first[k] = key[k] >= 0 ? __ldcg(p.cell_first + key[k]) : 0xFFFFFFFFu;
    // only the point that opens the pillar needs its size: a second, dependent gather for under half of the points costs
    // less than an unconditional one for all of them (the L2 serves ~40 scattered sectors per clock, whatever their size)
#pragma unroll
    for (int k = 0; k < kScanPer; ++k)
        cntm1[k] = (key[k] >= 0 && first[k] == static_cast<uint32_t>(i0 + k))
                       ? static_cast<uint32_t>(__ldcg(p.cell_row + key[k])) : 0u;
    unsigned long long val[kScanPer];
    unsigned flags = 0;
    unsigned long long tsum = 0;
#pragma unroll
    for (int k = 0; k < kScanPer; ++k) {
        val[k] = 0;
        // a tagged entry (list base already written by the pillar's owner) never equals a point index
        if (key[k] >= 0 && first[k] == static_cast<uint32_t>(i0 + k)) {
            val[k] = (1ull << 32) | static_cast<unsigned long long>(cntm1[k] + 1u);
            flags |= 1u << k;
        }
        tsum += val[k];
    }
    unsigned long long tile_sum;
    const unsigned long long thr_excl = block_scan_excl<kWarps>(tsum, s_warp, lane, warp, tile_sum);
    if (tid == 0) dbg_stamp(p.dbg, 9);   // gathers done, tile scanned
    if (tid == 0) st_relaxed_u64(&p.tile_agg[tile], pack_desc(tile_sum));  // visible to the successors at once

    // ---- per-frame tables from the insert kernel's pillar counts: first-appearance id and output row at each frame start
    // (off the critical path: the successors already have this tile's aggregate)
    {
        const uint32_t r0 = min(g0, mv), r1 = min(g1, mv);
        unsigned long long total;
        const unsigned long long ex = block_scan_excl<kWarps>(
            (static_cast<unsigned long long>(g0 + g1) << 32) | static_cast<unsigned long long>(r0 + r1), s_warp, lane, warp,
            total);
        const uint32_t eg = static_cast<uint32_t>(ex >> 32), er = static_cast<uint32_t>(ex & 0xFFFFFFFFull);
        if (f0 < p.nb) {
            s_gstart[f0] = eg;
            s_rowbase[f0] = er;
        }
        if (f0 + 1 < p.nb) {
            s_gstart[f0 + 1] = eg + g0;
            s_rowbase[f0 + 1] = er + r0;
        }
        if (tid == 0) {
            s_gstart[p.nb] = static_cast<uint32_t>(total >> 32);
            s_rowbase[p.nb] = static_cast<uint32_t>(total & 0xFFFFFFFFull);
        }
        if (tile == 0) {  // tile 0 publishes the tables for the later stages
            if (f0 < p.nb) {
                p.frame_gstart[f0] = eg;
                p.frame_rowbase[f0] = er;
                if (p.pillar_count) p.pillar_count[f0] = static_cast<int32_t>(r0);
            }
            if (f0 + 1 < p.nb) {
                p.frame_gstart[f0 + 1] = eg + g0;
                p.frame_rowbase[f0 + 1] = er + r0;
                if (p.pillar_count) p.pillar_count[f0 + 1] = static_cast<int32_t>(r1);
            }
            if (tid == 0) {
                p.frame_gstart[p.nb] = static_cast<uint32_t>(total >> 32);
                p.frame_rowbase[p.nb] = static_cast<uint32_t>(total & 0xFFFFFFFFull);
                if (p.pillar_count) p.pillar_count[p.nb] = static_cast<int32_t>(total & 0xFFFFFFFFull);
                p.hdr->total_pillars = static_cast<uint32_t>(total >> 32);
            }
        }
    }

    // look-back: aggregates of the tiles of my group that precede me (one parallel read) + prefix of the previous group
    // (a group is as many tiles as the CTA has threads: up to 512 K points resolve in ONE hop; the thread standing at the
    // tile's own position has no aggregate to read and fetches the previous group's prefix instead)
    const uint32_t group_first = tile & ~static_cast<uint32_t>(kDenseLookGroup - 1);
    unsigned long long look = 0;
    if (group_first + tid < tile) look = wait_desc(&p.tile_agg[group_first + tid]);
    else if (group_first + tid == tile && group_first > 0) look = wait_desc(&p.tile_prefix[group_first - 1]);
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) look += __shfl_xor_sync(kFullMask, look, s);
    if (lane == 0) s_look[warp] = look;
    __syncthreads();
    unsigned long long tile_excl = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) tile_excl += s_look[w];
    if (tid == 0 && (tile & (kDenseLookGroup - 1)) == kDenseLookGroup - 1)
        st_relaxed_u64(&p.tile_prefix[tile], pack_desc(tile_excl + tile_sum));

    if (tid == 0) dbg_stamp(p.dbg, 11);  // look-back resolved
    unsigned long long run = tile_excl + thr_excl;
    const uint32_t P = static_cast<uint32_t>(p.gd.max_points);
#pragma unroll
    for (int k = 0; k < kScanPer; ++k) {
        if (flags & (1u << k)) {
            const uint32_t g = static_cast<uint32_t>(run >> 32);
            const uint32_t base = static_cast<uint32_t>(run & 0xFFFFFFFFull);
            const uint32_t n = static_cast<uint32_t>(val[k] & 0xFFFFFFFFull);
            const uint32_t ukey = static_cast<uint32_t>(key[k]);
            const CellCoord c = decode_key(p, ukey);
            const uint32_t local = g - s_gstart[c.b];
            const int64_t row = static_cast<int64_t>(s_rowbase[c.b]) + local;
            const bool live = local < mv && row < p.capacity;
            p.cell_row[ukey] = live ? static_cast<int32_t>(row) : -1;
            p.cell_first[ukey] = kBaseTag | base;
            if (p.write_lists) {
                p.pillar_key[g] = ukey;
                p.pillar_list[g] = base;
                p.pillar_cnt[g] = n;
            }
            if (p.write_meta) {
                // indexed by the list START POSITION, so the feature kernel needs nothing but its own position to find it
                publish_pillar(p.pillar_meta, p.long_list, p.long_count, base, c, live ? static_cast<uint32_t>(row) : 0xFFFFFFFFu, n);
                if (live) {
                    if (p.voxel_coords)
                        *reinterpret_cast<int4 *>(p.voxel_coords + row * 4) = make_int4(
                            static_cast<int>(c.b), static_cast<int>(c.z), static_cast<int>(c.y), static_cast<int>(c.x));
                    if (p.voxel_num_points) p.voxel_num_points[row] = static_cast<int32_t>(min(n, P));
                }
            }
            run += val[k];
        }
    }
    if (tile == p.n_tiles - 1 && tid == 0)
        p.hdr->total_listed = static_cast<uint32_t>((tile_excl + tile_sum) & 0xFFFFFFFFull);
    if (tid == 0) dbg_stamp(p.dbg, 13);
}

__global__ void __launch_bounds__(kGroupThreads) k_place_dense(const __grid_constant__ PlaceParams p)
{
    if (threadIdx.x == 0) dbg_stamp(p.dbg, 14);
    // two points per thread.  Keys and arrival ranks were written by the insert kernel, which had completed before the
    // first CTA of the scan kernel passed ITS wait -- and this grid is launched only after all of them did: they are
    // loaded while the scan is still running, only the (key -> list base) gathers need the scan's results.
    const int64_t i0 = static_cast<int64_t>(blockIdx.x) * kInsTile + threadIdx.x;
    int32_t key[kInsPer];
    uint32_t arrival[kInsPer], base[kInsPer];
#pragma unroll
    for (int k = 0; k < kInsPer; ++k) {
        const int64_t i = i0 + k * kGroupThreads;
        key[k] = i < p.n ? __ldcg(p.point_slot + i) : -1;
        arrival[k] = i < p.n ? __ldcg(p.point_arrival + i) : 0u;
    }
    // ... and so are the records themselves (the points are an input of the call), all but their position
    float4 ra[kInsPer], rd[kInsPer];
    if (p.records) {
#pragma unroll
        for (int k = 0; k < kInsPer; ++k)
            if (key[k] >= 0)
                make_record(p, i0 + k * kGroupThreads, decode_key(p, static_cast<uint32_t>(key[k])), arrival[k], ra[k], rd[k]);
    }
    pdl_wait();
    pdl_trigger();
    if (threadIdx.x == 0) dbg_stamp(p.dbg, 16);
#pragma unroll
    for (int k = 0; k < kInsPer; ++k) base[k] = key[k] >= 0 ? (__ldcg(p.cell_first + key[k]) & ~kBaseTag) : 0u;
#pragma unroll
    for (int k = 0; k < kInsPer; ++k) {
        if (key[k] < 0) continue;
        const uint32_t pos = base[k] + arrival[k];
        if (p.sorted_idx) p.sorted_idx[pos] = static_cast<uint32_t>(i0 + k * kGroupThreads);
        if (p.records) {
            float4 *dst = reinterpret_cast<float4 *>(p.records + pos);
            dst[0] = ra[k];
            dst[1] = rd[k];
        }
    }
    if (threadIdx.x == 0) dbg_stamp(p.dbg, 17);
}

}  // namespace

cudaError_t launch_group_points_dense(const float *points, int64_t n, int stride, int col0, int c_point,
                                      const int32_t *frame_offsets, int nb, const GridDev &gd, const Workspace &ws,
                                      int32_t *pillar_count, bool want_index_lists, const PlaceExtras &extras,
                                      cudaStream_t st)
{
    cudaError_t err;
    // ONE fill per call: tile counter, per-frame pillar counters, scan descriptors and both planes of the cell table all
    // start from 0xFF bytes (counters count up from -1, a descriptor is ready once bit 63 is clear).
    {
        const size_t n_vec = ws.ff_bytes / 16;  // every piece of the region is 256-byte aligned
        // (a smaller grid, leaving SM room for the insert kernel's CTAs to stage their tiles meanwhile, measured the same)
        const unsigned fb = static_cast<unsigned>(tmin<size_t>((n_vec + 255) / 256, static_cast<size_t>(current_sm_count()) * 8));
        k_fill_ff<<<fb, 256, 0, st>>>(reinterpret_cast<uint4 *>(ws.ff_begin), n_vec, debug_times_ptr());
        note_launch();
    }
    if (n == 0) {
        if (pillar_count) {
            if ((err = cudaMemsetAsync(pillar_count, 0, sizeof(int32_t) * (nb + 1), st)) != cudaSuccess) return err;
            note_launch();
        }
        return cudaSuccess;
    }
    const size_t smem = sizeof(float) * kInsTile * stride;
    const int vec_ok = (reinterpret_cast<uintptr_t>(points) % 16 == 0) ? 1 : 0;  // tile starts are 512 rows apart
    const unsigned pb = static_cast<unsigned>((n + kInsTile - 1) / kInsTile);
    if ((err = launch_pdl(k_insert_dense, dim3(pb), dim3(kGroupThreads), smem, st, points, n, stride, col0, frame_offsets, nb,
                          gd, ws.cell_first, reinterpret_cast<uint32_t *>(ws.cell_row), ws.point_slot, ws.point_arrival,
                          ws.frame_new, vec_ok, ws.hdr, debug_times_ptr())) != cudaSuccess)
        return err;
    note_launch();

    const PlaceParams pp = make_place_params(points, n, stride, col0, c_point, gd, ws, want_index_lists, extras);
    ScanDenseParams sp{};
    sp.n = n;
    sp.n_tiles = ws.n_tiles;
    sp.point_key = ws.point_slot;
    sp.cell_first = ws.cell_first;
    sp.cell_row = ws.cell_row;
    sp.tile_agg = ws.tile_desc;
    sp.tile_prefix = ws.tile_prefix;
    sp.frame_new = ws.frame_new;
    sp.nb = nb;
    sp.hdr = ws.hdr;
    sp.frame_gstart = ws.frame_gstart;
    sp.frame_rowbase = ws.frame_rowbase;
    sp.pillar_count = pillar_count;
    sp.pillar_key = ws.pillar_key;
    sp.pillar_list = ws.pillar_list;
    sp.pillar_cnt = ws.pillar_cnt;
    sp.pillar_meta = ws.pillar_meta;
    sp.long_list = ws.long_list;
    sp.long_count = ws.long_count;
    sp.voxel_coords = pp.voxel_coords;
    sp.voxel_num_points = pp.voxel_num_points;
    sp.gd = gd;
    sp.sh_cells = pp.sh_cells;
    sp.sh_cells_xy = pp.sh_cells_xy;
    sp.sh_nx = pp.sh_nx;
    sp.capacity = extras.records ? extras.capacity : (1ll << 62);
    sp.write_lists = want_index_lists ? 1 : 0;
    sp.write_meta = extras.records ? 1 : 0;
    sp.dbg = debug_times_ptr();
    if ((err = launch_pdl(k_scan_dense, dim3(ws.n_tiles), dim3(kScanThreads), 0, st, sp)) != cudaSuccess) return err;
    note_launch();
    if (want_index_lists || extras.records) {
        if ((err = launch_pdl(k_place_dense, dim3(pb), dim3(kGroupThreads), 0, st, pp)) != cudaSuccess) return err;
        note_launch();
    }
    return cudaGetLastError();
}

}  // namespace pillars
