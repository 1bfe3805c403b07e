// asp.cu — activation-sparse SGEMV on a dense row-major A.
//
// Replaces asp_kernel_v0/v1/v2 (reference asp.cu:6-211).  The reference re-tiles A into
// 32x32 tiles (asp.cpp:3-14), gives each 32-column slab to a 4-warp CTA and tests x[j] != 0
// in the inner loop, one 128-byte tile row per warp load.  Here A stays row-major (a row is
// one contiguous N*4-byte run), and
//   * a CTA owns (512-column tile, row range); it first compacts its slice of x into a
//     shared-memory list of (row, x[row]) with x[row] != 0.0f (asp.cu:23's test, made a pass:
//     ballot + popc prefix, order preserving), so inactive rows are never addressed;
//   * every thread owns four adjacent output columns; a warp streams its 512 contiguous bytes
//     of every active row through a private cp.async ring in shared memory, kAspStages rows in
//     flight per warp (commit/wait groups: a true FIFO, no register or scoreboard limits), and
//     each lane reads back only the 16 bytes it copied itself, so no barrier is needed;
//   * row splits are summed in split order by the last CTA to arrive (integer ticket).
// Deterministic, no floating-point atomics.
#include <algorithm>

#include "common.cuh"
#include "plan.hpp"

namespace spmv {

namespace {

#ifndef SPMV_ASP_THREADS
#define SPMV_ASP_THREADS 128
#endif
constexpr int kAspThreads = SPMV_ASP_THREADS;
constexpr int kAspTile = kAspThreads * 4;     // output columns per CTA
constexpr int kAspChunk = 1024;               // rows compacted per pass
#ifndef SPMV_ASP_STAGES
#define SPMV_ASP_STAGES 16
#endif
constexpr int kAspStages = SPMV_ASP_STAGES; // rows in flight per warp

__global__ void __launch_bounds__(kAspThreads)
asp_kernel(const float *__restrict__ A, long long ld, const float *__restrict__ x, const YDst yd,
           float *__restrict__ partial, unsigned *__restrict__ tickets, int M, int N, int rows_per_split,
           int splits)
{
    __shared__ __align__(16) int rows_s[kAspChunk];
    __shared__ float xs_s[kAspChunk];
    __shared__ int wcnt[kAspThreads / 32];
    __shared__ int last_flag;
    extern __shared__ __align__(16) float4 ring_all[];    // (kAspThreads / 32) * kAspStages * 32

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x, split = blockIdx.y;
    const int c0 = tile * kAspTile + tid * 4;
    const bool col_ok = c0 < N;                           // N % 4 == 0: all four or none
    const int r_begin = split * rows_per_split;
    const int r_end = min(M, r_begin + rows_per_split);
    const float *Ac = A + c0;
    float4 *ring = ring_all + warp * kAspStages * 32;

    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r0 = r_begin; r0 < r_end; r0 += kAspChunk) {
        // ---- compaction of x[r0 .. r0+chunk): warp w takes a contiguous quarter ----------------
        constexpr int kSpan = kAspChunk / (kAspThreads / 32);   // 256 rows per warp
        constexpr int kSteps = kSpan / 32;                      // 8
        float xr[kSteps]; unsigned bal[kSteps];
        int cnt = 0;
#pragma unroll
        for (int k = 0; k < kSteps; k++) {
            const int row = r0 + warp * kSpan + k * 32 + lane;
            xr[k] = row < r_end ? __ldg(x + row) : 0.0f;
            bal[k] = __ballot_sync(kFull, xr[k] != 0.0f);
            cnt += __popc(bal[k]);
        }
        __syncthreads();                                  // previous chunk's list fully consumed
        if (lane == 0) wcnt[warp] = cnt;
        __syncthreads();
        int base = 0, total = 0;
#pragma unroll
        for (int w = 0; w < kAspThreads / 32; w++) {
            const int cw = wcnt[w];
            if (w < warp) base += cw;
            total += cw;
        }
        const unsigned lt = (1u << lane) - 1u;
#pragma unroll
        for (int k = 0; k < kSteps; k++) {
            if (xr[k] != 0.0f) {
                const int pos = base + __popc(bal[k] & lt);
                rows_s[pos] = r0 + warp * kSpan + k * 32 + lane;
                xs_s[pos] = xr[k];
            }
            base += __popc(bal[k]);
        }
        __syncthreads();

        // ---- stream the active rows ------------------------------------------------------------
        if (col_ok) {
            auto issue = [&](int i) {
                if (i < total)
                    cp_async16(ring + (i & (kAspStages - 1)) * 32 + lane,
                               Ac + (long long)rows_s[i] * ld);
                cp_async_commit();
            };
#pragma unroll
            for (int i = 0; i < kAspStages; i++) issue(i);
            for (int i = 0; i < total; i++) {
                cp_async_wait<kAspStages - 1>();          // row i has landed
                const float4 a = ring[(i & (kAspStages - 1)) * 32 + lane];
                const float xv = xs_s[i];
                acc.x = fmaf(a.x, xv, acc.x); acc.y = fmaf(a.y, xv, acc.y);
                acc.z = fmaf(a.z, xv, acc.z); acc.w = fmaf(a.w, xv, acc.w);
                issue(i + kAspStages);
            }
            cp_async_wait<0>();
        }
    }

    if (splits == 1) {
        if (col_ok) y_store4(yd, (size_t)c0 >> 2, acc);
        return;
    }
    const size_t npad = (size_t)gridDim.x * kAspTile;
    *reinterpret_cast<float4 *>(partial + (size_t)split * npad + (size_t)tile * kAspTile + tid * 4) = acc;
    const int n_valid = min(kAspTile, N - tile * kAspTile);
    __syncthreads();                                      // the row list is dead: reuse it as scratch
    split_reduce_finish(yd, partial, tickets, tile, splits, kAspTile, n_valid, npad, &last_flag,
                        reinterpret_cast<float4 *>(rows_s));
}

} // namespace

int launch_asp(spmv_plan *p, const float *d_x, const YDst &yd, cudaStream_t st)
{
    if (p->N == 0) return SPMV_OK;
    const int smem = (kAspThreads / 32) * kAspStages * 32 * (int)sizeof(float4);
    static int smem_set[16] = {0};
    if (smem > 48 * 1024 && p->device >= 0 && p->device < 16 && smem_set[p->device] < smem) {
        SPMV_CUDA(cudaFuncSetAttribute(asp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        smem_set[p->device] = smem;
    }
    asp_kernel<<<p->grid, kAspThreads, smem, st>>>(p->asp.A, (long long)p->asp.ld, d_x, yd, p->partial, p->tickets,
                                               (int)p->M, (int)p->N, p->asp.rows_per_split, p->row_splits);
    SPMV_CUDA(cudaGetLastError());
    return SPMV_OK;
}

// grid = (ceil(N/512), row splits): a little over two CTAs per SM, all resident at once — every
// CTA pays a fixed few microseconds (x compaction, first rows, split reduction), so more splits
// are slower (measured: 8-12 splits 24.7 us, 16 splits 26.9 us on config 2) — at least 64 rows each.
int configure_asp(spmv_plan *p, const spmv_options_t *o)
{
    p->block = kAspThreads;
    p->smem = (kAspThreads / 32) * kAspStages * 32 * (int)sizeof(float4);
    p->tile_width = kAspTile;
    p->col_tiles = (int)((p->N + kAspTile - 1) / kAspTile);
    p->asp.tile_cols = kAspTile;
    p->kernels_per_run = 1;
    const int64_t M = std::max<int64_t>(p->M, 1);
    int splits;
    if (o && o->row_splits > 0) splits = (int)std::min<int64_t>(o->row_splits, M);
    else splits = std::max(1, (9 * p->sm_count / 4 + p->col_tiles / 2) / std::max(1, p->col_tiles));
    int rps = (int)((M + splits - 1) / splits);
    if (!(o && o->row_splits > 0)) rps = std::max(64, rps);
    rps = (rps + 31) / 32 * 32;
    splits = (int)((M + rps - 1) / rps);
    p->asp.rows_per_split = rps;
    p->row_splits = splits;
    p->grid = dim3((unsigned)std::max(1, p->col_tiles), (unsigned)splits, 1);
    return alloc_split_scratch(p);
}

} // namespace spmv
