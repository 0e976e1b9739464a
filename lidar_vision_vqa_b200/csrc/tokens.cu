// BEV tokeniser: the head of the reference's VATLiDAR.forward (src/encoder-decoder/training/models/vat_lidar.py:206-253),
// the first consumer of the pillar canvas:
//
//     x = GELU(depthwise3x3(bev) + b)            :82-85,211      tokens = LayerNorm(x . Wp^T + bp)     :88-89,222-225
//     tokens += geo_mlp(x, y, r, sin, cos)       :92-97,229-231  tokens += view_embed[sector(cell)]    :101,245
//
// B200 shape of the problem.  The output [B, H*W, d] is d/C times the size of the canvas (4.3 GB for 16 frames of 512x512 at
// d = 256) and is the HBM stream that bounds the kernel; the 1x1 projection is a [cells x C] x [C x d] product.  A pillar
// canvas is ~95 % exact zeros, and a cell whose zero-padded 3x3 window holds only zeros produces a token that does not
// depend on the input:  LN(Wp . GELU(b) + bp) + PE(cell).  So:
//   * the positional table PE = geo_mlp(geom) + view_embed[sector] ([H*W, d]) and that background token are computed once
//     per (weights, H, W) by k_tok_pe / k_tok_background (eval mode: the reference recomputes geo_mlp every forward);
//   * k_bev_tokens reads the pillar ROWS (pillar_features + the BEV index map the scatter uses) instead of the dense
//     canvas, streams "background + PE" for cells with an empty window and runs the dense arithmetic only for the
//     others, on the FMA pipes in fp32 (packed FFMA2), G cells at a time per warp so that a projection column loaded once
//     feeds G cells;
//   * a CTA owns 32 consecutive cells of one canvas row for ALL frames, so PE is read once per cell, not once per frame;
//   * a dense canvas (the reference's own input format, forward(bev)) is first compacted into rows + index map by
//     k_canvas_to_rows (cells with any non-zero channel), then takes the same path.
#include "common.cuh"

namespace pillars {

namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kTokThreads = 256;
constexpr int kTokWarps = kTokThreads / 32;
constexpr int kTileX = 32;               // cells per CTA along x
constexpr int kCellsPerWarp = kTileX / kTokWarps;
constexpr int kFrameChunk = 16;          // frames whose index-map window is staged at once

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }

// packed fp32 pairs (FFMA2): w * (x, x) + c
__device__ __forceinline__ void fma2s(float &c0, float &c1, float w0, float w1, float x)
{
    unsigned long long rw, rx, rc;
    asm("mov.b64 %0, {%1, %2};" : "=l"(rw) : "f"(w0), "f"(w1));
    asm("mov.b64 %0, {%1, %1};" : "=l"(rx) : "f"(x));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rc) : "f"(c0), "f"(c1));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(rc) : "l"(rw), "l"(rx));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(c0), "=f"(c1) : "l"(rc));
}

struct TokParams {
    const float *feats;      // [rows, c] pillar rows
    const int32_t *cell_row; // [nb, h, w], -1 = empty cell
    int nb, h, w, c, d;
    const float *dw_w;       // [c, 9]   refine.0.weight
    const float *dw_b;       // [c]      refine.0.bias
    const float *wt;         // [c, d]   proj.weight transposed
    const float *pb;         // [d]      proj.bias
    const float *gamma, *beta;
    float eps;
    const float *pe;         // [h*w, d]
    const float *bg;         // [d]
    float *out;              // [nb, h*w, d]
};

// -------------------------------------------------------------------------------------------------------------------
// main kernel.  NQ = d / 128: a lane owns channels 4*(lane + 32*q) .. +3, q < NQ (512 contiguous bytes per warp and q).
// G = cells whose projection runs together in one warp: a projection column loaded once feeds G cells, but the G x NQ x 4
// accumulators set the register count and with it the warps in flight -- measured on B200 (16 x 512^2, d = 256): G = 8 at
// 2 CTAs/SM 1586 us, G = 4 at 3 CTAs/SM 1425 us; d = 512: G = 4 1475 us, G = 2 1645 us.
// dynamic shared memory: s_dw [10][c] (depthwise weights transposed + bias) | s_a [warps][c/4][G] float4 (refined activations)
// -------------------------------------------------------------------------------------------------------------------
template <int NQ, int G>
__global__ void __launch_bounds__(kTokThreads, (NQ * G <= 8) ? 3 : 2) k_bev_tokens(const __grid_constant__ TokParams p)
{
    extern __shared__ __align__(16) float s_dyn[];
    __shared__ int32_t s_map[kFrameChunk][3][kTileX + 2];
    __shared__ uint32_t s_act[kFrameChunk];                  // bit t: cell t of the tile has a non-empty window
    __shared__ uint16_t s_list[kTileX * kFrameChunk];        // (cell of the tile) * kFrameChunk + frame, CTA-wide
    __shared__ uint32_t s_count;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c = p.c, d = p.d, w = p.w, h = p.h;
    float *const s_dw = s_dyn;
    // activation tile of the warp's group: [quad][G] float4
    float *const s_a = s_dyn + 10 * c + warp * G * c;
    const int x0 = blockIdx.x * kTileX, y = blockIdx.y;

    for (int i = tid; i < 9 * c; i += kTokThreads) {
        const int ch = i / 9, k = i - ch * 9;
        s_dw[k * c + ch] = __ldg(p.dw_w + i);
    }
    for (int i = tid; i < c; i += kTokThreads) s_dw[9 * c + i] = __ldg(p.dw_b + i);

    // refine: lane = (cell g of the group, channel quad): consecutive lanes hold consecutive cells, so the activation tile
    // [quad][G] is written without bank conflicts and read back by the projection with one running pointer
    const int quads = c >> 2;
    const int rg = lane % G, rq = lane / G;
    constexpr int kQuadStep = 32 / G;

    for (int b0 = 0; b0 < p.nb; b0 += kFrameChunk) {
        const int nbb = min(kFrameChunk, p.nb - b0);
        __syncthreads();  // the previous chunk's readers are done
        if (tid == 0) s_count = 0;
        for (int i = tid; i < nbb * 3 * (kTileX + 2); i += kTokThreads) {
            const int bb = i / (3 * (kTileX + 2)), r = i - bb * 3 * (kTileX + 2);
            const int dy = r / (kTileX + 2), dx = r - dy * (kTileX + 2);
            const int yy = y + dy - 1, xx = x0 + dx - 1;
            int32_t v = -1;
            if (yy >= 0 && yy < h && xx >= 0 && xx < w)
                v = __ldg(p.cell_row + (static_cast<size_t>(b0 + bb) * h + yy) * w + xx);
            s_map[bb][dy][dx] = v;
        }
        __syncthreads();
        for (int i = tid; i < nbb * kTileX; i += kTokThreads) {  // consecutive lanes = consecutive cells of one frame
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

        // ---- pass A: stream the input-independent tokens, collect the (cell, frame) pairs that need arithmetic --------------
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
                    s_list[base + __popc(amask & ((1u << lane) - 1u))] = static_cast<uint16_t>(t * kFrameChunk + lane);
            }
            if (__popc(amask) == nbb) continue;
            float4 v[NQ];
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                const float4 a = __ldg(reinterpret_cast<const float4 *>(p.bg) + lane + 32 * q);
                const float4 e = __ldg(reinterpret_cast<const float4 *>(p.pe + cell * d) + lane + 32 * q);
                v[q] = make_float4(a.x + e.x, a.y + e.y, a.z + e.z, a.w + e.w);
            }
            float4 *dst = reinterpret_cast<float4 *>(p.out + (static_cast<size_t>(b0) * h * w + cell) * d) + lane;
            const size_t frame_step = static_cast<size_t>(h) * w * (d >> 2);
            for (int bb = 0; bb < nbb; ++bb, dst += frame_step) {
                if ((amask >> bb) & 1u) continue;
#pragma unroll
                for (int q = 0; q < NQ; ++q) __stcs(dst + 32 * q, v[q]);
            }
        }
        __syncthreads();
        const int n_list = static_cast<int>(s_count);

        // ---- pass B: G pairs at a time per warp, groups dealt round-robin over the CTA's warps ------------------------------
        for (int i0 = warp * G; i0 < n_list; i0 += kTokWarps * G) {
            const int cnt = min(G, n_list - i0);
            // refine: depthwise 3x3 over the window's pillar rows, + bias, GELU (vat_lidar.py:82-85)
            if (rg < cnt) {
                const int e = s_list[i0 + rg];
                const int t = e / kFrameChunk, bb = e - t * kFrameChunk;
                const float4 *rowp[9];
#pragma unroll
                for (int k = 0; k < 9; ++k) {
                    const int32_t r = s_map[bb][k / 3][t + k % 3];
                    rowp[k] = r >= 0 ? reinterpret_cast<const float4 *>(p.feats + static_cast<size_t>(r) * c) : nullptr;
                }
                for (int qd = rq; qd < quads; qd += kQuadStep) {
                    float4 f[9];
#pragma unroll
                    for (int k = 0; k < 9; ++k) f[k] = rowp[k] ? __ldg(rowp[k] + qd) : make_float4(0.f, 0.f, 0.f, 0.f);
                    float4 acc = *reinterpret_cast<const float4 *>(s_dw + 9 * c + 4 * qd);
#pragma unroll
                    for (int k = 0; k < 9; ++k) {
                        const float4 wk = *reinterpret_cast<const float4 *>(s_dw + k * c + 4 * qd);
                        acc.x = fmaf(wk.x, f[k].x, acc.x);
                        acc.y = fmaf(wk.y, f[k].y, acc.y);
                        acc.z = fmaf(wk.z, f[k].z, acc.z);
                        acc.w = fmaf(wk.w, f[k].w, acc.w);
                    }
                    const float4 act = make_float4(gelu_erf(acc.x), gelu_erf(acc.y), gelu_erf(acc.z), gelu_erf(acc.w));
                    reinterpret_cast<float4 *>(s_a)[qd * G + rg] = act;
                }
            }
            __syncwarp();
            // projection (vat_lidar.py:88,222): acc[g][q] += a[g][ch] * Wp^T[ch][lane's channels]
            float acc[G][NQ][4];
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                const float4 bq = __ldg(reinterpret_cast<const float4 *>(p.pb) + lane + 32 * q);
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    acc[g][q][0] = bq.x; acc[g][q][1] = bq.y; acc[g][q][2] = bq.z; acc[g][q][3] = bq.w;
                }
            }
            const float4 *wp = reinterpret_cast<const float4 *>(p.wt) + lane;
            const int d4 = d >> 2;
            const float4 *ap = reinterpret_cast<const float4 *>(s_a);          // [quad][G]: one running pointer, immediate offsets
            const float4 *const ap_end = ap + static_cast<size_t>(quads) * G;
            for (; ap != ap_end; ap += G) {
                float4 av[G];
#pragma unroll
                for (int g = 0; g < G; ++g) av[g] = ap[g];
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {
                    float4 wv[NQ];
#pragma unroll
                    for (int q = 0; q < NQ; ++q) wv[q] = __ldg(wp + 32 * q);
                    wp += d4;
#pragma unroll
                    for (int g = 0; g < G; ++g) {
                        const float a = cc == 0 ? av[g].x : cc == 1 ? av[g].y : cc == 2 ? av[g].z : av[g].w;
#pragma unroll
                        for (int q = 0; q < NQ; ++q) {
                            fma2s(acc[g][q][0], acc[g][q][1], wv[q].x, wv[q].y, a);
                            fma2s(acc[g][q][2], acc[g][q][3], wv[q].z, wv[q].w, a);
                        }
                    }
                }
            }
            // LayerNorm over d (two-pass, as ATen), + PE, store (vat_lidar.py:225,231,245)
#pragma unroll
            for (int g = 0; g < G; ++g) {
                if (g < cnt) {
                    float s = 0.f;
#pragma unroll
                    for (int q = 0; q < NQ; ++q) s += (acc[g][q][0] + acc[g][q][1]) + (acc[g][q][2] + acc[g][q][3]);
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(kFull, s, o);
                    const float mean = s / static_cast<float>(d);
                    float ss = 0.f;
#pragma unroll
                    for (int q = 0; q < NQ; ++q)
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float dlt = acc[g][q][j] - mean;
                            ss = fmaf(dlt, dlt, ss);
                        }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(kFull, ss, o);
                    const float rstd = 1.f / sqrtf(ss / static_cast<float>(d) + p.eps);
                    const int e = s_list[i0 + g];
                    const int t = e / kFrameChunk, bb = e - t * kFrameChunk;
                    const size_t cell = static_cast<size_t>(y) * w + x0 + t;
                    float4 *dst = reinterpret_cast<float4 *>(p.out + (static_cast<size_t>(b0 + bb) * h * w + cell) * d) + lane;
                    const float4 *pe4 = reinterpret_cast<const float4 *>(p.pe + cell * d) + lane;
#pragma unroll
                    for (int q = 0; q < NQ; ++q) {
                        const float4 ga = __ldg(reinterpret_cast<const float4 *>(p.gamma) + lane + 32 * q);
                        const float4 be = __ldg(reinterpret_cast<const float4 *>(p.beta) + lane + 32 * q);
                        const float4 pe = __ldg(pe4 + 32 * q);
                        float4 o4;
                        o4.x = fmaf((acc[g][q][0] - mean) * rstd, ga.x, be.x) + pe.x;
                        o4.y = fmaf((acc[g][q][1] - mean) * rstd, ga.y, be.y) + pe.y;
                        o4.z = fmaf((acc[g][q][2] - mean) * rstd, ga.z, be.z) + pe.z;
                        o4.w = fmaf((acc[g][q][3] - mean) * rstd, ga.w, be.w) + pe.w;
                        __stcs(dst + 32 * q, o4);
                    }
                }
            }
            __syncwarp();  // s_a is rewritten by the next group
        }
    }
}

// -------------------------------------------------------------------------------------------------------------------
// Token of a cell with an empty window, before PE: LN(Wp . GELU(refine.bias) + bp).  One CTA.
// -------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTokThreads) k_tok_background(const float *__restrict__ dw_b, const float *__restrict__ wt,
                                                                 const float *__restrict__ pb, const float *__restrict__ gamma,
                                                                 const float *__restrict__ beta, float eps, int c, int d,
                                                                 float *__restrict__ bg)
{
    extern __shared__ float s_act_bg[];  // [c]
    __shared__ float s_red[kTokWarps];
    __shared__ float s_stat;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < c; i += kTokThreads) s_act_bg[i] = gelu_erf(dw_b[i]);
    __syncthreads();
    float y[4];
    float part = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int j = tid + k * kTokThreads;
        y[k] = 0.f;
        if (j < d) {
            float acc = pb[j];
            for (int ch = 0; ch < c; ++ch) acc = fmaf(s_act_bg[ch], wt[static_cast<size_t>(ch) * d + j], acc);
            y[k] = acc;
            part += acc;
        }
    }
    auto block_sum = [&](float v) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
        __syncthreads();
        if (lane == 0) s_red[warp] = v;
        __syncthreads();
        if (tid == 0) {
            float t = 0.f;
            for (int i = 0; i < kTokWarps; ++i) t += s_red[i];
            s_stat = t;
        }
        __syncthreads();
        return s_stat;
    };
    const float mean = block_sum(part) / static_cast<float>(d);
    part = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (tid + k * kTokThreads < d) part = fmaf(y[k] - mean, y[k] - mean, part);
    const float rstd = 1.f / sqrtf(block_sum(part) / static_cast<float>(d) + eps);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int j = tid + k * kTokThreads;
        if (j < d) bg[j] = fmaf((y[k] - mean) * rstd, gamma[j], beta[j]);
    }
}

// -------------------------------------------------------------------------------------------------------------------
// PE[cell] = W2 . GELU(W1 . geom[cell] + b1) + b2 + view_embed[sid[cell]]   (vat_lidar.py:92-97,229-231,245).  Once per
// (weights, H, W); 8 cells per CTA, thread j owns output channels j, j + 256, ...
// -------------------------------------------------------------------------------------------------------------------
constexpr int kPeCells = 8;
__global__ void __launch_bounds__(kTokThreads) k_tok_pe(const float *__restrict__ geom, const int32_t *__restrict__ sid,
                                                         int64_t cells, int d, const float *__restrict__ w1,
                                                         const float *__restrict__ b1, const float *__restrict__ w2t,
                                                         const float *__restrict__ b2, const float *__restrict__ view,
                                                         float *__restrict__ pe)
{
    extern __shared__ float s_hid[];  // [kPeCells][d]
    const int tid = threadIdx.x;
    const int64_t cell0 = static_cast<int64_t>(blockIdx.x) * kPeCells;
    for (int i = tid; i < kPeCells * d; i += kTokThreads) {
        const int g = i / d, k = i - g * d;
        const int64_t cell = cell0 + g;
        float v = 0.f;
        if (cell < cells) {
            v = b1[k];
#pragma unroll
            for (int a = 0; a < 5; ++a) v = fmaf(w1[k * 5 + a], geom[cell * 5 + a], v);
            v = gelu_erf(v);
        }
        s_hid[i] = v;
    }
    __syncthreads();
    for (int j = tid; j < d; j += kTokThreads) {
        float acc[kPeCells];
#pragma unroll
        for (int g = 0; g < kPeCells; ++g) acc[g] = b2[j];
        for (int k = 0; k < d; ++k) {
            const float wv = __ldg(w2t + static_cast<size_t>(k) * d + j);
#pragma unroll
            for (int g = 0; g < kPeCells; ++g) acc[g] = fmaf(wv, s_hid[g * d + k], acc[g]);
        }
#pragma unroll
        for (int g = 0; g < kPeCells; ++g) {
            const int64_t cell = cell0 + g;
            if (cell < cells) pe[cell * d + j] = acc[g] + view[static_cast<size_t>(sid[cell]) * d + j];
        }
    }
}

// -------------------------------------------------------------------------------------------------------------------
// Dense canvas [B, C, H, W] -> rows [n, C] + index map: a cell with any non-zero channel gets a row (order irrelevant:
// rows are storage, the map is what the tokeniser follows).  One warp per 32 consecutive cells; channels in chunks of 64
// transposed through shared memory so that both the plane reads and the row writes are coalesced.
// -------------------------------------------------------------------------------------------------------------------
constexpr int kRowsWarps = 4;
__global__ void __launch_bounds__(32 * kRowsWarps) k_canvas_to_rows(const float *__restrict__ bev, int nb, int c, int64_t plane,
                                                                   int32_t *__restrict__ cell_row, float *__restrict__ rows,
                                                                   uint32_t *__restrict__ counter)
{
    __shared__ float s_t[kRowsWarps][32][65];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t groups_per_frame = (plane + 31) / 32;
    const int64_t grp = static_cast<int64_t>(blockIdx.x) * kRowsWarps + warp;
    if (grp >= groups_per_frame * nb) return;
    const int b = static_cast<int>(grp / groups_per_frame);
    const int64_t cell0 = (grp - b * groups_per_frame) * 32, cell = cell0 + lane;
    const bool in = cell < plane;
    const float *src = bev + static_cast<size_t>(b) * c * plane + cell;
    bool any = false;
    if (in)
        for (int ch = 0; ch < c; ++ch) any |= __ldg(src + static_cast<size_t>(ch) * plane) != 0.f;
    const unsigned m = __ballot_sync(kFull, any);
    uint32_t base = 0;
    if (lane == 0 && m) base = atomicAdd(counter, static_cast<uint32_t>(__popc(m)));
    base = __shfl_sync(kFull, base, 0);
    const int32_t row = any ? static_cast<int32_t>(base + __popc(m & ((1u << lane) - 1u))) : -1;
    if (in) cell_row[static_cast<size_t>(b) * plane + cell] = row;
    if (!m) return;
    for (int ch0 = 0; ch0 < c; ch0 += 64) {
        const int nch = min(64, c - ch0);
        __syncwarp();
        for (int k = 0; k < nch; ++k) s_t[warp][lane][k] = in ? __ldg(src + static_cast<size_t>(ch0 + k) * plane) : 0.f;
        __syncwarp();
        for (int l = 0; l < 32; ++l) {
            if (!((m >> l) & 1u)) continue;
            const int32_t r = __shfl_sync(kFull, row, l);
            for (int k = lane; k < nch; k += 32) rows[static_cast<size_t>(r) * c + ch0 + k] = s_t[warp][l][k];
        }
    }
}

// Vectorised variant (plane % 4 == 0, 16-byte aligned canvas): a lane scans 4 consecutive cells with 16-byte loads, a warp
// 128 cells, eight channel planes in flight per lane.  Occupied cells of a sparse group are then fetched one by one (a cell's
// channels are 4-byte gathers from planes that were just read, i.e. L2 hits, written as one coalesced row); a group with
// more than a quarter of its cells occupied goes through the shared-memory transpose in 32-cell sub-groups instead.
__global__ void __launch_bounds__(32 * kRowsWarps) k_canvas_to_rows_v4(const float *__restrict__ bev, int nb, int c, int64_t plane,
                                                                      int32_t *__restrict__ cell_row, float *__restrict__ rows,
                                                                      uint32_t *__restrict__ counter)
{
    __shared__ float s_t[kRowsWarps][32][65];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t groups_per_frame = (plane + 127) / 128;
    const int64_t grp = static_cast<int64_t>(blockIdx.x) * kRowsWarps + warp;
    if (grp >= groups_per_frame * nb) return;
    const int b = static_cast<int>(grp / groups_per_frame);
    const int64_t cell0 = (grp - b * groups_per_frame) * 128, cell = cell0 + 4 * lane;
    const bool in = cell < plane;  // plane % 4 == 0: a lane's four cells are inside or outside together
    const float *fbase = bev + static_cast<size_t>(b) * c * plane;
    uint32_t bx = 0, by = 0, bz = 0, bw = 0;  // OR of the value bits without the sign: non-zero <=> some channel != +-0
    if (in) {
        const float4 *src = reinterpret_cast<const float4 *>(fbase + cell);
        const size_t step = static_cast<size_t>(plane >> 2);
        int ch = 0;
        for (; ch + 8 <= c; ch += 8) {
            float4 v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = __ldg(src + (ch + k) * step);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                bx |= __float_as_uint(v[k].x); by |= __float_as_uint(v[k].y);
                bz |= __float_as_uint(v[k].z); bw |= __float_as_uint(v[k].w);
            }
        }
        for (; ch < c; ++ch) {
            const float4 v = __ldg(src + ch * step);
            bx |= __float_as_uint(v.x); by |= __float_as_uint(v.y); bz |= __float_as_uint(v.z); bw |= __float_as_uint(v.w);
        }
    }
    const bool any[4] = {(bx & 0x7fffffffu) != 0u, (by & 0x7fffffffu) != 0u, (bz & 0x7fffffffu) != 0u, (bw & 0x7fffffffu) != 0u};
    unsigned m[4];
    int total = 0, before[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        m[j] = __ballot_sync(kFull, any[j]);
        before[j] = total;
        total += __popc(m[j]);
    }
    uint32_t base = 0;
    if (lane == 0 && total) base = atomicAdd(counter, static_cast<uint32_t>(total));
    base = __shfl_sync(kFull, base, 0);
    int32_t row[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
        row[j] = any[j] ? static_cast<int32_t>(base + before[j] + __popc(m[j] & ((1u << lane) - 1u))) : -1;
    if (in) *reinterpret_cast<int4 *>(cell_row + static_cast<size_t>(b) * plane + cell) = make_int4(row[0], row[1], row[2], row[3]);
    if (!total) return;
    if (total <= 32) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            unsigned mm = m[j];
            while (mm) {
                const int l = __ffs(mm) - 1;
                mm &= mm - 1;
                const int32_t r = __shfl_sync(kFull, row[j], l);
                const float *cp = fbase + cell0 + 4 * l + j;
                for (int k = lane; k < c; k += 32) rows[static_cast<size_t>(r) * c + k] = __ldg(cp + static_cast<size_t>(k) * plane);
            }
        }
        return;
    }
    // dense group: 32-cell sub-groups through the transpose tile; sub-group s holds cells cell0 + 32 s + lane', owned by
    // lane (8 s + lane' / 4), component lane' % 4
    for (int sgrp = 0; sgrp < 4; ++sgrp) {
        const int owner = 8 * sgrp + (lane >> 2), comp = lane & 3;
        int32_t my_row = -1;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int32_t rj = __shfl_sync(kFull, row[j], owner);
            if (comp == j) my_row = rj;
        }
        const unsigned sm = __ballot_sync(kFull, my_row >= 0);
        if (!sm) continue;
        const int64_t scell = cell0 + 32 * sgrp + lane;
        const bool sin = scell < plane;
        for (int ch0 = 0; ch0 < c; ch0 += 64) {
            const int nch = min(64, c - ch0);
            __syncwarp();
            for (int k = 0; k < nch; ++k) s_t[warp][lane][k] = sin ? __ldg(fbase + static_cast<size_t>(ch0 + k) * plane + scell) : 0.f;
            __syncwarp();
            for (int l = 0; l < 32; ++l) {
                if (!((sm >> l) & 1u)) continue;
                const int32_t r = __shfl_sync(kFull, my_row, l);
                for (int k = lane; k < nch; k += 32) rows[static_cast<size_t>(r) * c + ch0 + k] = s_t[warp][l][k];
            }
        }
    }
}

template <int NQ, int G>
cudaError_t launch_tokens_t(const TokParams &p, cudaStream_t st)
{
    const size_t smem = sizeof(float) * (10 * static_cast<size_t>(p.c) + static_cast<size_t>(kTokWarps) * G * p.c);
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_bev_tokens<NQ, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return e;
    }
    const dim3 grid(static_cast<unsigned>((p.w + kTileX - 1) / kTileX), static_cast<unsigned>(p.h));
    k_bev_tokens<NQ, G><<<grid, kTokThreads, smem, st>>>(p);
    note_launch();
    return cudaGetLastError();
}

}  // namespace

bool tokens_shape_supported(int c, int d) { return c >= 4 && c % 4 == 0 && c <= 512 && d >= 128 && d % 128 == 0 && d <= 1024; }

cudaError_t launch_tokens_prepare(const TokenizerDev &tk, const float *geom, const int32_t *sid, int h, int w, const float *w1,
                                  const float *b1, const float *w2t, const float *b2, const float *view, float *pe, float *bg,
                                  cudaStream_t st)
{
    k_tok_background<<<1, kTokThreads, sizeof(float) * tk.c, st>>>(tk.dw_b, tk.wt, tk.pb, tk.gamma, tk.beta, tk.eps, tk.c,
                                                                   tk.d, bg);
    note_launch();
    const int64_t cells = static_cast<int64_t>(h) * w;
    if (cells > 0) {
        const unsigned blocks = static_cast<unsigned>((cells + kPeCells - 1) / kPeCells);
        k_tok_pe<<<blocks, kTokThreads, sizeof(float) * kPeCells * tk.d, st>>>(geom, sid, cells, tk.d, w1, b1, w2t, b2, view, pe);
        note_launch();
    }
    return cudaGetLastError();
}

cudaError_t launch_canvas_to_rows(const float *bev, int nb, int c, int h, int w, int32_t *cell_row, float *rows,
                                  uint32_t *counter, cudaStream_t st)
{
    cudaError_t e = cudaMemsetAsync(counter, 0, sizeof(uint32_t), st);
    note_launch();
    if (e != cudaSuccess) return e;
    const int64_t plane = static_cast<int64_t>(h) * w;
    if (plane % 4 == 0 && reinterpret_cast<uintptr_t>(bev) % 16 == 0 && reinterpret_cast<uintptr_t>(cell_row) % 16 == 0) {
        const int64_t groups4 = (plane + 127) / 128 * nb;
        if (groups4 == 0) return cudaSuccess;
        k_canvas_to_rows_v4<<<static_cast<unsigned>((groups4 + kRowsWarps - 1) / kRowsWarps), 32 * kRowsWarps, 0, st>>>(
            bev, nb, c, plane, cell_row, rows, counter);
        note_launch();
        return cudaGetLastError();
    }
    const int64_t groups = (plane + 31) / 32 * nb;
    if (groups == 0) return cudaSuccess;
    k_canvas_to_rows<<<static_cast<unsigned>((groups + kRowsWarps - 1) / kRowsWarps), 32 * kRowsWarps, 0, st>>>(
        bev, nb, c, plane, cell_row, rows, counter);
    note_launch();
    return cudaGetLastError();
}

cudaError_t launch_bev_tokens(const TokenizerDev &tk, const float *feats, const int32_t *cell_row, int nb, int h, int w,
                              float *out, cudaStream_t st)
{
    if (nb == 0 || h == 0 || w == 0) return cudaSuccess;
    TokParams p{};
    p.feats = feats; p.cell_row = cell_row; p.nb = nb; p.h = h; p.w = w; p.c = tk.c; p.d = tk.d;
    p.dw_w = tk.dw_w; p.dw_b = tk.dw_b; p.wt = tk.wt; p.pb = tk.pb; p.gamma = tk.gamma; p.beta = tk.beta; p.eps = tk.eps;
    p.pe = tk.pe; p.bg = tk.bg; p.out = out;
    switch (tk.d / 128) {
        case 1: return launch_tokens_t<1, 8>(p, st);
        case 2: return launch_tokens_t<2, 4>(p, st);
        case 3: return launch_tokens_t<3, 4>(p, st);
        case 4: return launch_tokens_t<4, 4>(p, st);
        case 5: return launch_tokens_t<5, 2>(p, st);
        case 6: return launch_tokens_t<6, 2>(p, st);
        case 7: return launch_tokens_t<7, 2>(p, st);
        case 8: return launch_tokens_t<8, 2>(p, st);
    }
    return cudaErrorInvalidValue;
}

}  // namespace pillars
