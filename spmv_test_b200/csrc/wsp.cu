// wsp.cu — weight-sparse SGEMV on the CSR(A^T) "V4" streams (formats.hpp: HostWsp).
//
// Replaces wsp_kernel_v0/v1 (reference wsp.cu:4-138) and csr_naive_kernel
// (csr_naive.cu:6-23).  The reference assigns one warp per output column and walks a column
// bitmap: bitmap word -> popc -> address -> one 4-byte load per lane, a dependent chain with
// one load in flight per lane.  Here every output column is a contiguous, 16-byte aligned
// run of (float4 values, 4 x row index) groups, so a team of T threads streams it with
// 128-bit coalesced loads, several per thread in flight, gathers x from shared memory
// (staged once per CTA by a 1-D bulk async copy), and reduces with a fixed shuffle tree.
// T is chosen per length bin (row-length binning), so a power-law matrix (BASELINE config 4)
// gives 4-thread teams to its short rows and whole CTAs to its long ones.
//
// Determinism: lane partials (4 chains, fixed association) -> xor-shuffle tree -> fixed-order
// cross-warp sum.  No atomics.
#include <algorithm>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "common.cuh"
#include "plan.hpp"

namespace spmv {

namespace {

constexpr int kWspBlock = 256;
constexpr int kWspUnroll = 4;
#ifndef SPMV_WSP_MERGED_CTAS
#define SPMV_WSP_MERGED_CTAS 4
#endif
constexpr int kWspMergedCtas = SPMV_WSP_MERGED_CTAS;      // resident CTAs per SM of the L2-gather kernels (register budget and bin grids)

template <typename IdxVec> struct IdxTraits;
template <> struct IdxTraits<uint2> {   // 4 x u16
    static __device__ __forceinline__ uint2 load(const uint2 *p) { return ldg_stream_u2(p); }
    static __device__ __forceinline__ void unpack(const uint2 &v, uint32_t (&i)[4])
    {
        i[0] = v.x & 0xffffu; i[1] = v.x >> 16; i[2] = v.y & 0xffffu; i[3] = v.y >> 16;
    }
    static __device__ __forceinline__ uint2 pad(uint32_t M) { uint32_t w = M | (M << 16); return make_uint2(w, w); }
};
template <> struct IdxTraits<uint4> {   // 4 x u32
    static __device__ __forceinline__ uint4 load(const uint4 *p) { return ldg_stream_u4(p); }
    static __device__ __forceinline__ void unpack(const uint4 &v, uint32_t (&i)[4])
    {
        i[0] = v.x; i[1] = v.y; i[2] = v.z; i[3] = v.w;
    }
    static __device__ __forceinline__ uint4 pad(uint32_t M) { return make_uint4(M, M, M, M); }
};

// T threads cooperate on one output column.  XS: x lives in shared memory.  (bid, nblocks):
// this CTA's position among the CTAs working on the same column list.
template <typename IdxVec, int T, bool XS>
__device__ __forceinline__ void
wsp_body(const float4 *__restrict__ vals, const IdxVec *__restrict__ idx,
         const uint32_t *__restrict__ colptr, const int32_t *__restrict__ cols, int ncols,
         const float *__restrict__ x, const YDst &yd, uint32_t M, int x_bulk_ok, int bid, int nblocks)
{
    extern __shared__ __align__(16) float xs[];          // M + 1 (+pad) floats when XS
    __shared__ __align__(8) uint64_t bar;
    __shared__ float red[kWspBlock / 32];

    const int tid = threadIdx.x;
    if (XS) {
        const uint32_t m4 = M & ~3u;                      // bulk part: whole 16-byte units
        if (x_bulk_ok && m4) {
            if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
            __syncthreads();
            if (tid == 0) {
                uint32_t done = 0;
                mbar_expect_tx(&bar, m4 * 4u);
                while (done < m4 * 4u) {                  // <= 32 KB per bulk copy
                    uint32_t n = min(m4 * 4u - done, 32768u);
                    bulk_g2s(reinterpret_cast<char *>(xs) + done, reinterpret_cast<const char *>(x) + done, n, &bar);
                    done += n;
                }
            }
            for (uint32_t j = m4 + tid; j < M; j += kWspBlock) xs[j] = x[j];
        } else {
            for (uint32_t j = tid; j < M; j += kWspBlock) xs[j] = x[j];
        }
        if (tid < 4) xs[M + tid] = 0.0f;                  // the pad slot (and alignment slack)
    }

    constexpr int kTeams = kWspBlock / T;
    const int team_local = tid / T;
    const int tl = tid % T;
    const int lane = tid & 31;
    const int team = bid * kTeams + team_local;
    const int nteams = nblocks * kTeams;

    // Issue the first column's loads before waiting for x: the A stream does not depend on it.
    if (XS) {
        if (x_bulk_ok && (M & ~3u)) mbar_wait(&bar, 0);
        __syncthreads();
    }

    // Sub-warp teams share a warp and the shuffles below name all 32 lanes, so every team of a warp makes the same
    // number of trips: a team past the end of the list walks an empty column and stores nothing.
    for (int k = team;; k += nteams) {
        const bool act = k < ncols;
        if (T < 32 ? !__any_sync(kFull, act) : !act) break;
        const int c = act ? (cols ? cols[k] : k) : 0;
        uint32_t g0 = 0u, g1 = 0u;
        if (act) { g0 = colptr[c]; g1 = colptr[c + 1]; }
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        for (uint32_t g = g0 + tl; g < g1; g += T * kWspUnroll) {
            float4 v[kWspUnroll];
            IdxVec iv[kWspUnroll];
#pragma unroll
            for (int u = 0; u < kWspUnroll; u++) {
                const uint32_t gg = g + u * T;
                if (gg < g1) { v[u] = ldg_stream_f4(vals + gg); iv[u] = IdxTraits<IdxVec>::load(idx + gg); }
                else { v[u] = make_float4(0.f, 0.f, 0.f, 0.f); iv[u] = IdxTraits<IdxVec>::pad(XS ? M : 0u); }
            }
#pragma unroll
            for (int u = 0; u < kWspUnroll; u++) {
                uint32_t i[4];
                IdxTraits<IdxVec>::unpack(iv[u], i);
                float x0, x1, x2, x3;
                if (XS) { x0 = xs[i[0]]; x1 = xs[i[1]]; x2 = xs[i[2]]; x3 = xs[i[3]]; }
                else {
                    x0 = i[0] < M ? __ldg(x + i[0]) : 0.f; x1 = i[1] < M ? __ldg(x + i[1]) : 0.f;
                    x2 = i[2] < M ? __ldg(x + i[2]) : 0.f; x3 = i[3] < M ? __ldg(x + i[3]) : 0.f;
                }
                a0 = fmaf(v[u].x, x0, a0); a1 = fmaf(v[u].y, x1, a1);
                a2 = fmaf(v[u].z, x2, a2); a3 = fmaf(v[u].w, x3, a3);
            }
        }
        float acc = (a0 + a1) + (a2 + a3);
        if (T >= 32) {
            acc = warp_sum(acc);
            if (T == 32) {
                if (lane == 0) y_store(yd, c, acc);
            } else {
                constexpr int W = T / 32;                 // warps per team
                const int wt = (tid / 32) % (W > 0 ? W : 1);
                const int bar_id = 1 + team_local;        // named barrier per team (0 = __syncthreads)
                if (lane == 0) red[tid / 32] = acc;
                asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(T) : "memory");
                if (wt == 0 && lane == 0) {
                    float s = 0.f;
                    for (int w = 0; w < W; w++) s += red[team_local * W + w];
                    y_store(yd, c, s);
                }
                asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(T) : "memory");
            }
        } else {
#pragma unroll
            for (int s = T / 2; s >= 1; s >>= 1) acc += __shfl_xor_sync(kFull, acc, s);
            if (tl == 0 && act) y_store(yd, c, acc);
        }
    }
}

// Short columns of a skewed matrix with x gathered through L2 (BASELINE config 4: a 4-thread team, one
// group per thread, mean 16 non-zeros per column).  One column per team at a time is a chain of four
// dependent memory round trips (bin list -> column range -> values / row ids -> x) with one load per
// thread in flight at each step: ncu showed 74 % of the stall samples on the long scoreboard at 45 %
// occupancy (profiles/r02_ncu_c4_wsp_merged.txt).  Here a team takes C columns at once and issues each
// step's loads for all of them before it uses any.  Per column the arithmetic and its order are those of
// wsp_body (same per-thread group sequence, same 4 chains, same shuffle tree): bit-identical results.
// Same-box A/B on config 4, us per call (resident CTAs per SM x columns per team; bin grids of 4 / 8 / 16 CTAs per
// SM): 8 x 1 (round 1) 102.4 / 107.0 / 112.9; 4 x 2 94.1 / 98.2 / 102.9; 6 x 2 100.5 / 104.9 / 111.6; 3 x 4 121.3 /
// 107.4 / 124.3; 2 x 8 141.4 / 157.9 / 173.9.  Two columns at a time with half the threads is the optimum: beyond
// that the gathers are limited by the L2 sector rate (16 M sectors for 16 M useful words), not by latency.
#ifndef SPMV_WSP_SHORT_COLS
#define SPMV_WSP_SHORT_COLS 2
#endif
template <typename IdxVec, int T, int C>
__device__ __forceinline__ void
wsp_body_short(const float4 *__restrict__ vals, const IdxVec *__restrict__ idx, const uint32_t *__restrict__ colptr,
               const int32_t *__restrict__ cols, int ncols, const float *__restrict__ x, const YDst &yd, uint32_t M,
               int bid, int nblocks)
{
    static_assert(T <= 32, "a team is part of one warp");
    constexpr int kTeams = kWspBlock / T;
    const int tid = threadIdx.x, tl = tid % T;
    const int team = bid * kTeams + tid / T, nteams = nblocks * kTeams;
    for (int k0 = team * C;; k0 += nteams * C) {            // (same trip count for every team of a warp: see wsp_body)
        if (T < 32 ? !__any_sync(kFull, k0 < ncols) : !(k0 < ncols)) break;
        int c[C];
        uint32_t g0[C], g1[C];
#pragma unroll
        for (int j = 0; j < C; j++) c[j] = k0 + j < ncols ? (cols ? __ldg(cols + k0 + j) : k0 + j) : -1;
#pragma unroll
        for (int j = 0; j < C; j++) {
            g0[j] = 0u; g1[j] = 0u;
            if (c[j] >= 0) { g0[j] = __ldg(colptr + c[j]); g1[j] = __ldg(colptr + c[j] + 1); }
        }
        float4 v[C];
        IdxVec iv[C];
#pragma unroll
        for (int j = 0; j < C; j++) {                       // every column's first group of this thread
            const uint32_t gg = g0[j] + tl;
            if (gg < g1[j]) { v[j] = ldg_stream_f4(vals + gg); iv[j] = IdxTraits<IdxVec>::load(idx + gg); }
            else { v[j] = make_float4(0.f, 0.f, 0.f, 0.f); iv[j] = IdxTraits<IdxVec>::pad(0u); }
        }
        float xg[C][4];
#pragma unroll
        for (int j = 0; j < C; j++) {
            uint32_t i[4];
            IdxTraits<IdxVec>::unpack(iv[j], i);
#pragma unroll
            for (int q = 0; q < 4; q++) xg[j][q] = i[q] < M ? __ldg(x + i[q]) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < C; j++) {
            float a0 = fmaf(v[j].x, xg[j][0], 0.f), a1 = fmaf(v[j].y, xg[j][1], 0.f);
            float a2 = fmaf(v[j].z, xg[j][2], 0.f), a3 = fmaf(v[j].w, xg[j][3], 0.f);
            for (uint32_t g = g0[j] + tl + T; g < g1[j]; g += T) {   // (columns of this bin hold at most 8 T groups)
                const float4 vv = ldg_stream_f4(vals + g);
                uint32_t i[4];
                IdxTraits<IdxVec>::unpack(IdxTraits<IdxVec>::load(idx + g), i);
                const float x0 = i[0] < M ? __ldg(x + i[0]) : 0.f, x1 = i[1] < M ? __ldg(x + i[1]) : 0.f;
                const float x2 = i[2] < M ? __ldg(x + i[2]) : 0.f, x3 = i[3] < M ? __ldg(x + i[3]) : 0.f;
                a0 = fmaf(vv.x, x0, a0); a1 = fmaf(vv.y, x1, a1); a2 = fmaf(vv.z, x2, a2); a3 = fmaf(vv.w, x3, a3);
            }
            float acc = (a0 + a1) + (a2 + a3);
            if (T == 32) {
                acc = warp_sum(acc);
            } else {
#pragma unroll
                for (int sft = T / 2; sft >= 1; sft >>= 1) acc += __shfl_xor_sync(kFull, acc, sft);
            }
            if (tl == 0 && c[j] >= 0) y_store(yd, c[j], acc);
        }
    }
}

template <typename IdxVec, int T, bool XS>
__global__ void __launch_bounds__(kWspBlock)
wsp_kernel(const float4 *__restrict__ vals, const IdxVec *__restrict__ idx,
           const uint32_t *__restrict__ colptr, const int32_t *__restrict__ cols, int ncols,
           const float *__restrict__ x, const YDst yd, uint32_t M, int x_bulk_ok)
{
    pdl_wait();
    wsp_body<IdxVec, T, XS>(vals, idx, colptr, cols, ncols, x, yd, M, x_bulk_ok, blockIdx.x, gridDim.x);
}

// All length bins of a skewed matrix in ONE launch (x gathered through L2): a CTA finds its bin
// from its block index and runs that bin's team size.  Seven launches with seven ramps and tails
// become one (BASELINE config 4).
constexpr int kMaxBins = 8;
struct BinTable {
    int n;
    int first_cta[kMaxBins + 1];
    int T[kMaxBins];
    int ncols[kMaxBins];
    const int32_t *cols[kMaxBins];
};

template <typename IdxVec>
__global__ void __launch_bounds__(kWspBlock, SPMV_WSP_MERGED_CTAS)
wsp_merged_kernel(const float4 *__restrict__ vals, const IdxVec *__restrict__ idx,
                  const uint32_t *__restrict__ colptr, const BinTable tb, const float *__restrict__ x,
                  const YDst yd, uint32_t M)
{
    pdl_wait();
    int b = 0;
    while (b + 1 < tb.n && (int)blockIdx.x >= tb.first_cta[b + 1]) b++;
    const int bid = blockIdx.x - tb.first_cta[b], nb = tb.first_cta[b + 1] - tb.first_cta[b];
    const int32_t *cols = tb.cols[b];
    const int ncols = tb.ncols[b];
    switch (tb.T[b]) {
    case 4: wsp_body_short<IdxVec, 4, SPMV_WSP_SHORT_COLS>(vals, idx, colptr, cols, ncols, x, yd, M, bid, nb); break;
    case 8: wsp_body_short<IdxVec, 8, SPMV_WSP_SHORT_COLS>(vals, idx, colptr, cols, ncols, x, yd, M, bid, nb); break;
    case 16: wsp_body_short<IdxVec, 16, 2>(vals, idx, colptr, cols, ncols, x, yd, M, bid, nb); break;
    case 32: wsp_body<IdxVec, 32, false>(vals, idx, colptr, cols, ncols, x, yd, M, 0, bid, nb); break;
    case 64: wsp_body<IdxVec, 64, false>(vals, idx, colptr, cols, ncols, x, yd, M, 0, bid, nb); break;
    case 128: wsp_body<IdxVec, 128, false>(vals, idx, colptr, cols, ncols, x, yd, M, 0, bid, nb); break;
    default: wsp_body<IdxVec, 256, false>(vals, idx, colptr, cols, ncols, x, yd, M, 0, bid, nb); break;
    }
}

// ---- long columns: warp per column, cp.async ring ------------------------------------------
// A warp walks its columns as one flat sequence of chunks (32 groups = 128 non-zeros); chunk
// loads go through a private shared-memory ring (kRingStages commit groups in flight: a true
// FIFO, independent of register and scoreboard limits), partial chunks are zero-filled,
// x is gathered from shared memory.  The packer
// deals each chunk's entries so that the 32 lanes' gathers fall into distinct banks.
// Round 2 A/Bs of how a chunk reaches the ring (same box, us per call, config 2 / 0 / 3; profiles/r02_notes.md):
//   cp.async, 16 + 8 bytes per lane (this code)                          22.4 / 13.9 / 22.0
//   two 1-D bulk async copies per chunk (TMA engine, cp.async.bulk +
//   one mbarrier per ring slot, issued by lane 0; the patch is kept as
//   profiles/r02_wsp_bulk_ring_variant.patch)                           30.1 / 17.6 / 33.3   (batch of 4: 85.8 vs 44.6)
//   chunks in flight in registers (8 x 6 registers, three CTAs per SM)   21.7 / 13.9 / 26.4
// The bulk engine loses on 768-byte pieces (as it did on 256..1024-byte row segments in round 1's
// tools/ubench/bulk_vs_ldgsts.cu), and this kernel's LSU work is dominated by the x gathers, not the staging.
#ifndef SPMV_WSP_STAGES
#define SPMV_WSP_STAGES 8
#endif
constexpr int kRingStages = SPMV_WSP_STAGES;
constexpr int kRingWarps = 8;

template <typename IdxVec> struct RingCopy;
template <> struct RingCopy<uint2> {
    static __device__ __forceinline__ void copy(uint2 *d, const uint2 *s) { cp_async8(d, s); }
};
template <> struct RingCopy<uint4> {
    static __device__ __forceinline__ void copy(uint4 *d, const uint4 *s) { cp_async16(d, s); }
};

// B > 1: batched (multi-vector) form — B activation vectors x[b] (row stride ldx) against the
// same A: every value / index pair fetched from HBM is used B times.  Each
// vector's arithmetic is exactly that of the single-vector kernel, so results are bit-identical
// to B separate calls.  Batched mode is single-panel, single-destination.
template <typename IdxVec, int B>
__global__ void __launch_bounds__(kRingWarps * 32)
wsp_ring_kernel(const float4 *__restrict__ vals, const IdxVec *__restrict__ idx,
                const uint32_t *__restrict__ colptr, const int32_t *__restrict__ cols, int ncols,
                const float *__restrict__ x, const YDst yd, uint32_t M_total, int x_bulk_ok, int xs_bytes,
                uint32_t panel_rows, int n_total, float *__restrict__ partial, long long ldx, long long ldy)
{
    extern __shared__ __align__(16) unsigned char wsm[];
    __shared__ __align__(8) uint64_t bar;
    float *xs = reinterpret_cast<float *>(wsm);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // row panel of this CTA (blockIdx.y): its slice of x goes to shared memory, its lists start at
    // colptr[panel * N]; rows past the slice (last panel) and the pad slot read as zero
    const uint32_t panel = blockIdx.y;
    const uint32_t M = min(panel_rows, M_total - panel * panel_rows);
    x += (size_t)panel * panel_rows;
    colptr += (size_t)panel * n_total;
    float4 *ring_v = reinterpret_cast<float4 *>(wsm + xs_bytes) + warp * kRingStages * 32;
    const int wpc = blockDim.x >> 5;                      // warps per CTA: 4 .. kRingWarps, chosen per plan (configure_wsp)
    IdxVec *ring_i = reinterpret_cast<IdxVec *>(wsm + xs_bytes + wpc * kRingStages * 32 * 16) + warp * kRingStages * 32;

    pdl_wait();
    // x -> shared memory (1-D bulk async copies when aligned), overlapped with the first chunks
    const uint32_t xstride = (uint32_t)xs_bytes / (4u * B);   // floats per vector in shared memory
    const uint32_t m4 = M & ~3u;
    const bool bulk = x_bulk_ok && m4;
    if (bulk) {
        if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
        __syncthreads();
        if (tid == 0) {
            mbar_expect_tx(&bar, m4 * 4u * B);
            for (int b = 0; b < B; b++)
                for (uint32_t done = 0; done < m4 * 4u; done += 32768u)
                    bulk_g2s(reinterpret_cast<char *>(xs + b * xstride) + done,
                             reinterpret_cast<const char *>(x + b * ldx) + done, min(m4 * 4u - done, 32768u), &bar);
        }
        for (int b = 0; b < B; b++)
            for (uint32_t j = m4 + tid; j < M; j += blockDim.x) xs[b * xstride + j] = x[b * ldx + j];
    } else {
        for (int b = 0; b < B; b++)
            for (uint32_t j = tid; j < M; j += blockDim.x) xs[b * xstride + j] = x[b * ldx + j];
    }
    for (int b = 0; b < B; b++)
        for (uint32_t j = M + tid; j < panel_rows + 4; j += blockDim.x) xs[b * xstride + j] = 0.0f;   // tail of the last panel + pad slot

    // ---- flat iterator over (column, chunk) ---------------------------------------------------------
    // This warp's columns are positions kbase, kbase + nwarps, ... of the column list.  Their
    // (column id, first group, end group) are fetched 32 at a time — lane l holds the l-th
    // upcoming column — one batch ahead of use, so even one-chunk columns keep the ring full.
    const int nwarps = gridDim.x * wpc;
    const int kbase = blockIdx.x * wpc + warp;
    int bc = -1, nc = -1; uint32_t b0 = 0, b1 = 0, n0 = 0, n1 = 0;   // current / next batch (per lane)
    int jb = 0;                                            // index of the next column inside the current batch
    int batch0 = 0;                                        // list position (in this warp's sequence) of the next batch
    auto fetch_batch = [&](int j0) {
        const long long kk = (long long)kbase + (long long)(j0 + lane) * nwarps;
        nc = -1; n0 = 0; n1 = 0;
        if (kk < ncols) { nc = cols ? cols[kk] : (int)kk; n0 = __ldg(colptr + nc); n1 = __ldg(colptr + nc + 1); }
    };
    fetch_batch(0);
    bc = nc; b0 = n0; b1 = n1;
    batch0 = 32;
    fetch_batch(batch0);
    uint32_t g = 0, gend = 0; int ccol = -1; bool live = true;
    auto next_col = [&]() {
        if (jb == 32) {                                    // promote the prefetched batch
            bc = nc; b0 = n0; b1 = n1; jb = 0;
            batch0 += 32;
            fetch_batch(batch0);
        }
        ccol = __shfl_sync(kFull, bc, jb);
        g = __shfl_sync(kFull, b0, jb);
        gend = __shfl_sync(kFull, b1, jb);
        jb++;
        if (ccol < 0) live = false;
    };
    next_col();

    int fin[kRingStages];                                  // >= 0: chunk ends column `fin`; -1: no; -2: empty stage
    auto issue = [&](int s) {
        fin[s] = -2;
        if (live) {
            const uint32_t gg = g + lane;
            const bool ok = gg < gend;
            // lanes past the column's end take value 0 (zero-fill) and the row ids of the
            // chunk's first group: real rows of this column, so 0 * x[row] adds nothing new
            const uint32_t gs = ok ? gg : g;
            cp_async16_zfill(ring_v + s * 32 + lane, vals + gs, ok);
            RingCopy<IdxVec>::copy(ring_i + s * 32 + lane, idx + gs);
            g += 32;
            fin[s] = -1;
            if (g >= gend) { fin[s] = ccol; next_col(); }
        }
        cp_async_commit();
    };
#pragma unroll
    for (int s = 0; s < kRingStages; s++) issue(s);

    if (bulk) mbar_wait(&bar, 0);
    __syncthreads();                                      // x is in shared memory

    float a0[B], a1[B], a2[B], a3[B];
#pragma unroll
    for (int b = 0; b < B; b++) a0[b] = a1[b] = a2[b] = a3[b] = 0.f;
    while (fin[0] != -2) {
#pragma unroll
        for (int s = 0; s < kRingStages; s++) {
            cp_async_wait<kRingStages - 1>();
            if (fin[s] != -2) {
                const float4 v = ring_v[s * 32 + lane];
                uint32_t i[4];
                IdxTraits<IdxVec>::unpack(ring_i[s * 32 + lane], i);
#pragma unroll
                for (int b = 0; b < B; b++) {
                    const float *xb = xs + b * xstride;
                    a0[b] = fmaf(v.x, xb[i[0]], a0[b]); a1[b] = fmaf(v.y, xb[i[1]], a1[b]);
                    a2[b] = fmaf(v.z, xb[i[2]], a2[b]); a3[b] = fmaf(v.w, xb[i[3]], a3[b]);
                }
                if (fin[s] >= 0) {
#pragma unroll
                    for (int b = 0; b < B; b++) {
                        const float t = warp_sum((a0[b] + a1[b]) + (a2[b] + a3[b]));
                        if (lane == 0) {
                            if (gridDim.y == 1) y_store(yd, (size_t)b * ldy + fin[s], t);
                            else partial[(size_t)panel * n_total + fin[s]] = t;
                        }
                        a0[b] = a1[b] = a2[b] = a3[b] = 0.f;
                    }
                }
            }
            issue(s);
        }
    }
    cp_async_wait<0>();
}

// y = sum over the row panels, in panel order
__global__ void __launch_bounds__(256)
wsp_combine_kernel(const float *__restrict__ partial, int panels, int n, const YDst yd)
{
    pdl_wait();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    float s = partial[c];
    for (int p = 1; p < panels; p++) s += partial[(size_t)p * n + c];
    y_store(yd, c, s);
}

struct Bin { int T; std::vector<int32_t> cols; };

} // namespace

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
struct WspBinDev { int T; int32_t *cols; int ncols; int grid; bool ring; int smem; int warps; };   // warps: per CTA of the ring kernel

struct WspState {            // hangs off the plan through plan->wsp_state
    std::vector<WspBinDev> bins;
    int panels = 1;
    int64_t panel_rows = 0;
    float *partial = nullptr;    // [panels][N] when panels > 1
};

template <typename IdxVec, int T, bool XS>
static int launch_one(const spmv_plan *p, const WspBinDev &b, const float *x, const YDst &y, cudaStream_t st,
                      size_t smem, int x_bulk_ok)
{
    auto k = wsp_kernel<IdxVec, T, XS>;
    static int smem_set[16] = {0};
    if (smem > 48 * 1024 && p->device >= 0 && p->device < 16 && smem_set[p->device] < (int)smem) {
        SPMV_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set[p->device] = (int)smem;
    }
    SPMV_CUDA(launch_k(k, dim3(b.grid), dim3(kWspBlock), smem, st, reinterpret_cast<const float4 *>(p->wsp.vals),
                       reinterpret_cast<const IdxVec *>(p->wsp.idx), p->wsp.colptr, b.cols, b.ncols, x, y, (uint32_t)p->M,
                       x_bulk_ok));
    return SPMV_OK;
}

template <typename IdxVec, bool XS>
static int launch_T(const spmv_plan *p, const WspBinDev &b, const float *x, const YDst &y, cudaStream_t st,
                    size_t smem, int ok)
{
    switch (b.T) {
    case 4: return launch_one<IdxVec, 4, XS>(p, b, x, y, st, smem, ok);
    case 8: return launch_one<IdxVec, 8, XS>(p, b, x, y, st, smem, ok);
    case 16: return launch_one<IdxVec, 16, XS>(p, b, x, y, st, smem, ok);
    case 32: return launch_one<IdxVec, 32, XS>(p, b, x, y, st, smem, ok);
    case 64: return launch_one<IdxVec, 64, XS>(p, b, x, y, st, smem, ok);
    case 128: return launch_one<IdxVec, 128, XS>(p, b, x, y, st, smem, ok);
    case 256: return launch_one<IdxVec, 256, XS>(p, b, x, y, st, smem, ok);
    }
    return set_error(SPMV_ERR_ARG, "wsp: bad team size %d", b.T);
}

template <typename IdxVec, int B>
static int launch_ring(const spmv_plan *p, const WspBinDev &b, const float *x, const YDst &y, cudaStream_t st, int ok,
                       long long ldx, long long ldy)
{
    auto k = wsp_ring_kernel<IdxVec, B>;
    const int xs_bytes = p->smem * B;
    const int smem = b.smem + p->smem * (B - 1);
    static int smem_set[16] = {0};
    if (smem > 48 * 1024 && p->device >= 0 && p->device < 16 && smem_set[p->device] < smem) {
        SPMV_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        smem_set[p->device] = smem;
    }
    const WspState *ws = reinterpret_cast<const WspState *>(p->wsp_state);
    SPMV_CUDA(launch_k(k, dim3(b.grid, ws->panels), dim3((b.warps > 0 ? b.warps : kRingWarps) * 32), smem, st, reinterpret_cast<const float4 *>(p->wsp.vals),
                       reinterpret_cast<const IdxVec *>(p->wsp.idx), p->wsp.colptr, b.cols, b.ncols, x, y, (uint32_t)p->M, ok,
                       xs_bytes, (uint32_t)ws->panel_rows, (int)p->N, ws->partial, ldx, ldy));
    if (ws->panels > 1)
        SPMV_CUDA(launch_k(wsp_combine_kernel, dim3((unsigned)((p->N + 255) / 256)), dim3(256), 0, st, ws->partial, ws->panels,
                           (int)p->N, y));
    return SPMV_OK;
}

// Batched form: B in {2, 4} vectors in one pass over A when the plan is a single ring bin with one
// row panel and the B copies of x fit shared memory; returns SPMV_ERR_UNSUPPORTED otherwise (the
// caller then runs the vectors one by one).
int launch_wsp_batch(spmv_plan *p, const float *d_x, long long ldx, const YDst &d_y, long long ldy, int B, cudaStream_t st)
{
    const WspState *s = reinterpret_cast<const WspState *>(p->wsp_state);
    if (!s || s->bins.size() != 1 || !s->bins[0].ring || s->panels != 1 || (B != 2 && B != 4)) return SPMV_ERR_UNSUPPORTED;
    const WspBinDev &b = s->bins[0];
    const int smem = b.smem + p->smem * (B - 1);
    if (smem > (p->max_smem_optin > 0 ? p->max_smem_optin : 227 * 1024)) return SPMV_ERR_UNSUPPORTED;
    const int ok = ((reinterpret_cast<uintptr_t>(d_x) & 15) == 0 && ldx % 4 == 0) ? 1 : 0;
    if (p->wsp.index_bits == 16)
        return B == 2 ? launch_ring<uint2, 2>(p, b, d_x, d_y, st, ok, ldx, ldy) : launch_ring<uint2, 4>(p, b, d_x, d_y, st, ok, ldx, ldy);
    return B == 2 ? launch_ring<uint4, 2>(p, b, d_x, d_y, st, ok, ldx, ldy) : launch_ring<uint4, 4>(p, b, d_x, d_y, st, ok, ldx, ldy);
}

int launch_wsp(spmv_plan *p, const float *d_x, const YDst &d_y, cudaStream_t st)
{
    if (p->N == 0) return SPMV_OK;
    const WspState *s = reinterpret_cast<const WspState *>(p->wsp_state);
    const int ok = ((reinterpret_cast<uintptr_t>(d_x) & 15) == 0) ? 1 : 0;
    if (!p->wsp.x_in_smem && s->bins.size() > 1 && (int)s->bins.size() <= kMaxBins) {
        BinTable tb{};
        tb.n = (int)s->bins.size();
        int total = 0;
        for (int i = 0; i < tb.n; i++) {
            const WspBinDev &b = s->bins[(size_t)i];
            tb.first_cta[i] = total; tb.T[i] = b.T; tb.ncols[i] = b.ncols; tb.cols[i] = b.cols;
            total += b.grid;
        }
        tb.first_cta[tb.n] = total;
        if (p->wsp.index_bits == 16)
            SPMV_CUDA(launch_k(wsp_merged_kernel<uint2>, dim3(total), dim3(kWspBlock), 0, st, reinterpret_cast<const float4 *>(p->wsp.vals),
                               reinterpret_cast<const uint2 *>(p->wsp.idx), p->wsp.colptr, tb, d_x, d_y, (uint32_t)p->M));
        else
            SPMV_CUDA(launch_k(wsp_merged_kernel<uint4>, dim3(total), dim3(kWspBlock), 0, st, reinterpret_cast<const float4 *>(p->wsp.vals),
                               reinterpret_cast<const uint4 *>(p->wsp.idx), p->wsp.colptr, tb, d_x, d_y, (uint32_t)p->M));
        return SPMV_OK;
    }
    for (const WspBinDev &b : s->bins) {
        int rc;
        if (b.ring) {
            if (p->wsp.index_bits == 16) rc = launch_ring<uint2, 1>(p, b, d_x, d_y, st, ok, 0, 0);
            else rc = launch_ring<uint4, 1>(p, b, d_x, d_y, st, ok, 0, 0);
            if (rc) return rc;
            continue;
        }
        const size_t smem = p->wsp.x_in_smem ? (size_t)p->smem : 0;
        if (p->wsp.index_bits == 16)
            rc = p->wsp.x_in_smem ? launch_T<uint2, true>(p, b, d_x, d_y, st, smem, ok)
                                  : launch_T<uint2, false>(p, b, d_x, d_y, st, smem, ok);
        else
            rc = p->wsp.x_in_smem ? launch_T<uint4, true>(p, b, d_x, d_y, st, smem, ok)
                                  : launch_T<uint4, false>(p, b, d_x, d_y, st, smem, ok);
        if (rc) return rc;
    }
    return SPMV_OK;
}

static int pow2_ceil(int64_t v) { int p = 1; while (p < v) p <<= 1; return p; }

int configure_wsp(spmv_plan *p, const HostWsp &w, const spmv_options_t *o)
{
    WspState *s = new WspState();
    p->wsp_state = s;
    p->wsp.index_bits = w.index_bits;
    s->panels = w.panels;
    p->wsp.panels = w.panels;
    p->wsp.panel_rows = w.panels > 1 ? w.panel_rows : w.M;
    s->panel_rows = w.panels > 1 ? w.panel_rows : w.M;
    if (w.panels > 1) {
        SPMV_CUDA(cudaMalloc(&s->partial, (size_t)w.panels * std::max<int64_t>(w.N, 1) * sizeof(float)));
        p->scratch_bytes += (int64_t)w.panels * w.N * 4;
    }
    // x (one row panel of it) in shared memory when it (plus the pad slot) fits comfortably next
    // to several CTAs/SM
    const size_t xbytes = ((size_t)s->panel_rows + 4) * sizeof(float);
    p->wsp.x_in_smem = xbytes <= 96 * 1024 && w.index_bits == 16;
    if (w.index_bits == 32 && xbytes <= 96 * 1024) p->wsp.x_in_smem = true;
    p->smem = p->wsp.x_in_smem ? (int)((xbytes + 15) & ~(size_t)15) : 0;
    p->block = kWspBlock;

    // ---- row-length binning --------------------------------------------------------------
    // target ~4 groups (16 non-zeros) per thread; bins are powers of two in [4, 256].
    const int64_t N = w.N;
    // (round 2: 4 instead of 8 — one unrolled iteration of kWspUnroll loads per thread, i.e. ONE memory round trip per
    // column instead of two: config 1 6.72 -> 5.13 us, config 4 94.9 -> 89.3 us, configs 0 / 2 unchanged)
    int gpt = kWspUnroll;
    if (const char *e = std::getenv("SPMV_WSP_GROUPS_PER_THREAD")) gpt = std::max(1, std::atoi(e));   // development knob
    auto team_for = [&](int64_t groups) {
        int t = pow2_ceil((groups + gpt - 1) / gpt);
        return std::min(256, std::max(4, t));
    };
    int64_t gmin = INT64_MAX, gmax = 0;
    for (int64_t i = 0; i < N * w.panels; i++) {
        int64_t g = (int64_t)w.colptr[i + 1] - w.colptr[i];
        gmin = std::min(gmin, g); gmax = std::max(gmax, g);
    }
    const int64_t gmean = N ? (w.groups + N * w.panels - 1) / (N * w.panels) : 0;
    int forced = 0;
    if (o && o->warps_per_col > 0) forced = std::min(256, 32 * pow2_ceil(o->warps_per_col));
    std::vector<Bin> bins;
    if (forced || N == 0 || w.panels > 1 || team_for(gmax) <= 2 * team_for(std::max<int64_t>(gmin, 1))) {
        bins.push_back({forced ? forced : team_for(gmean), {}});           // one bin: all columns, no list
    } else {
        const int Ts[7] = {4, 8, 16, 32, 64, 128, 256};
        for (int t : Ts) bins.push_back({t, {}});
        for (int64_t i = 0; i < N; i++) {
            int t = team_for((int64_t)w.colptr[i + 1] - w.colptr[i]);
            for (Bin &b : bins) if (b.T == t) b.cols.push_back((int32_t)i);
        }
        std::vector<Bin> keep;
        for (Bin &b : bins) if (!b.cols.empty()) keep.push_back(std::move(b));
        bins.swap(keep);
    }
    // occupancy-sized persistent grids
    const bool merged = !p->wsp.x_in_smem && bins.size() > 1 && (int)bins.size() <= kMaxBins;   // (launch_wsp's condition)
    int ctas_per_sm = p->wsp.x_in_smem ? std::max(1, std::min(8, (int)((200 * 1024) / std::max(p->smem, 1)))) : merged ? kWspMergedCtas : 8;
    if (const char *e = std::getenv("SPMV_WSP_BIN_CTAS")) ctas_per_sm = std::max(1, std::atoi(e));   // development knob
    for (Bin &b : bins) {
        WspBinDev d{};
        d.T = b.T;
        d.ncols = b.cols.empty() ? (int)N : (int)b.cols.size();
        d.cols = nullptr;
        if (!b.cols.empty()) {
            SPMV_CUDA(cudaMalloc(&d.cols, b.cols.size() * sizeof(int32_t)));
            SPMV_CUDA(cudaMemcpy(d.cols, b.cols.data(), b.cols.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
            p->device_bytes += (int64_t)b.cols.size() * 4;
        }
        const int teams_per_cta = kWspBlock / b.T;
        const int64_t need = ((int64_t)d.ncols + teams_per_cta - 1) / teams_per_cta;
        d.grid = (int)std::max<int64_t>(1, std::min<int64_t>(need, (int64_t)p->sm_count * ctas_per_sm));
        d.ring = false; d.smem = 0;
        if (p->wsp.x_in_smem && (b.T >= 32 || w.panels > 1) && !(o && o->warps_per_col > 0 && w.panels == 1)) {
            // long columns: warp per column through the cp.async ring
            d.ring = true;
            const int chunk_bytes = 32 * (16 + (w.index_bits == 16 ? 8 : 16));
            auto ring_smem = [&](int warps) { return p->smem + warps * kRingStages * chunk_bytes; };
            auto resident_of = [&](int warps) { return std::max(1, std::min(8, (220 * 1024) / (ring_smem(warps) + 1024))); };
            d.warps = kRingWarps;
            d.smem = ring_smem(kRingWarps);
            const int resident = resident_of(kRingWarps);
            const int64_t want = ((int64_t)d.ncols + kRingWarps - 1) / kRingWarps;
            d.grid = (int)std::max<int64_t>(1, std::min<int64_t>(want, ((int64_t)p->sm_count * resident + w.panels - 1) / w.panels));
            // Whole columns are dealt to the warps, so with more columns than resident warps the kernel lasts
            // ceil(columns / warps) columns while the average warp has columns / warps of them: config 2's 14336
            // columns on 3552 warps (three 8-warp CTAs per SM) are 4.04 per warp — 128 warps stream a fifth column
            // alone at the end (0.81 of the rate); config 0's 4096 on 3552 are 1.15 -> 2 (0.58).  Choose the
            // geometry (CTAs per SM, 4..8 warps per CTA; one resident wave) that wastes least, preferring more warps.
            // Same-box A/B, us per call, config 2 / 0 / 3: three 8-warp CTAs per SM (round 1) 22.36 / 13.02 / 22.69;
            // 3 x 7 22.18 / 12.96 / 22.17; 3 x 6 21.41 / 12.64 / 24.05; 2 x 8 21.35 / 12.07 / 21.49; 2 x 7 (what this
            // rule picks for all three) 21.15 / 11.65 / 21.18.  A 12-deep ring instead of 8 is slower everywhere.
            if (w.panels == 1 && d.ncols > p->sm_count * resident * kRingWarps) {
                double best = 0.0;
                for (int wp = kRingWarps; wp >= 4; wp--) {
                    for (int c = resident_of(wp); c >= 1; c--) {
                        const int64_t wt = (int64_t)p->sm_count * c * wp;
                        if (wt < (int64_t)p->sm_count * 12) continue;         // keep at least 12 warps per SM streaming
                        const int64_t cpw = (d.ncols + wt - 1) / wt;
                        const double eff = (double)d.ncols / ((double)wt * (double)cpw);
                        if (eff > best + 0.02) { best = eff; d.warps = wp; d.grid = p->sm_count * c; d.smem = ring_smem(wp); }
                    }
                }
            }
            if (const char *e = std::getenv("SPMV_WSP_RING_GEOM")) {      // development knob: "ctas_per_sm,warps_per_cta"
                int c = 0, wp = 0;
                if (std::sscanf(e, "%d,%d", &c, &wp) == 2 && c >= 1 && wp >= 1 && wp <= kRingWarps) {
                    d.warps = wp; d.smem = ring_smem(wp);
                    d.grid = (int)std::min<int64_t>(((int64_t)d.ncols + wp - 1) / wp, (int64_t)p->sm_count * c);
                }
            }
        }
        s->bins.push_back(d);
    }
    p->kernels_per_run = (int)s->bins.size() + (w.panels > 1 ? 1 : 0);
    // merged launch.  (Round 2: sharing ONE resident wave of sm_count * 8 CTAs among the bins in proportion to
    // their thread-iterations, instead of a full wave per bin, was much slower on config 4 — 160 vs 107 us: the
    // 2.3 waves of the per-bin grids are what balances bins of very different per-column cost.)
    if (!p->wsp.x_in_smem && s->bins.size() > 1 && (int)s->bins.size() <= kMaxBins) p->kernels_per_run = 1;
    p->grid = dim3(s->bins.empty() ? 1 : s->bins[0].grid, 1, 1);
    p->wsp.warps_per_col = s->bins.empty() ? 0 : std::max(1, s->bins[0].T / 32);
    p->wsp_team = s->bins.empty() ? 0 : s->bins[0].T;
    return SPMV_OK;
}

int clone_wsp_state(const spmv_plan *src, spmv_plan *dst)
{
    const WspState *s = reinterpret_cast<const WspState *>(src->wsp_state);
    WspState *d = new WspState();
    dst->wsp_state = d;
    if (!s) return SPMV_OK;
    d->panels = s->panels; d->panel_rows = s->panel_rows;
    if (s->partial) SPMV_CUDA(cudaMalloc(&d->partial, (size_t)s->panels * std::max<int64_t>(src->N, 1) * sizeof(float)));
    for (const WspBinDev &b : s->bins) {
        WspBinDev nb = b;
        nb.cols = nullptr;
        if (b.cols) {
            SPMV_CUDA(cudaMalloc(&nb.cols, (size_t)b.ncols * sizeof(int32_t)));
            d->bins.push_back(nb);
            SPMV_CUDA(cudaMemcpy(nb.cols, b.cols, (size_t)b.ncols * sizeof(int32_t), cudaMemcpyDeviceToDevice));
        } else {
            d->bins.push_back(nb);
        }
    }
    return SPMV_OK;
}

void destroy_wsp_state(spmv_plan *p)
{
    WspState *s = reinterpret_cast<WspState *>(p->wsp_state);
    if (!s) return;
    for (WspBinDev &b : s->bins) if (b.cols) cudaFree(b.cols);
    if (s->partial) cudaFree(s->partial);
    delete s;
    p->wsp_state = nullptr;
}

} // namespace spmv
