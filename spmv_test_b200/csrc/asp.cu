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
//     of every active row — long lists with kAspRegs rows in flight in registers (plain 128-bit
//     loads), short lists through a private cp.async ring in shared memory, kAspStages rows in
//     flight per warp (commit/wait groups: a true FIFO, no register or scoreboard limits; each
//     lane reads back only the 16 bytes it copied itself, so no barrier is needed);
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
constexpr int kAspMaxBatch = 4;               // vectors per batched pass
constexpr int kAspChunk = 1024;               // rows compacted per pass
#ifndef SPMV_ASP_STAGES
#define SPMV_ASP_STAGES 16
#endif
constexpr int kAspStages = SPMV_ASP_STAGES; // rows in flight per warp
static_assert((kAspStages & (kAspStages - 1)) == 0, "the ring is indexed with a mask");
// > 0: the single-vector kernel keeps that many rows in flight per warp in REGISTERS (one 128-bit load per lane and
// row, issued where it is written) instead of the cp.async ring: a ring costs LSU wavefronts twice (copy-in and
// read-back, profiles/r02_notes.md), and 16 rows x 512 B x ~6 warps per SM is only ~50 KB in flight.
// Same-box A/B, us per call on config 2 / 0 / 3 (ring = 16-deep cp.async ring for every chunk): ring 23.69 / 11.56 /
// 9.31; registers for chunks with >= 96 active rows (plans whose row ranges hold 192 rows): 32 rows in flight 22.41 / 11.53 / 9.35, 40: 22.00 / 11.59 / 9.43,
// 48: 21.76 / 11.55 / 9.46; 48 rows for chunks with >= 48 active rows (kept): 21.74 / 11.54 / 9.22 against 21.73 /
// 11.70 / 9.57 on that box; a 32-deep ring for the short lists: 21.79 / 12.03 / 10.00.  With 16 rows in flight registers LOSE to the ring (28.8 us): a warp has six scoreboards,
// so waiting for the oldest of 16 loads also waits for younger ones that share its scoreboard, while cp.async groups
// are an exact FIFO; the register path wins by depth (48 x 512 B per warp).  Short lists (config 3: ~45 active rows
// per CTA) stay on the ring, whose ramp is cheaper.
#ifndef SPMV_ASP_REGS
#define SPMV_ASP_REGS 48
#endif
constexpr int kAspRegs = SPMV_ASP_REGS;
#ifndef SPMV_ASP_REGS_MIN
#define SPMV_ASP_REGS_MIN 48
#endif
constexpr int kAspRegsMin = SPMV_ASP_REGS_MIN;   // active rows in a CTA's chunk from which the register path is taken (below: the ring)

__device__ __forceinline__ float4 asp_row_load(const float *p, bool ok)
{
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ok) asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}

// B > 1: batched form — B activation vectors x[b] (row stride ldx) against the
// same A: a row is streamed once if ANY vector is active there and used for all of them (a vector
// with x[b][row] == 0 adds an exact zero), so every y[b] is bit-identical to a single-vector call.
// R: rows in flight in registers for long lists (0: the ring only — the instance plans with short row ranges get,
// whose register count then stays small).
template <int B, int R>
__global__ void __launch_bounds__(kAspThreads)
asp_kernel(const float *__restrict__ A, long long ld, const float *__restrict__ x, const YDst yd,
           float *__restrict__ partial, unsigned *__restrict__ tickets, int M, int N, int rows_per_split,
           int splits, long long ldx, long long ldy)
{
    __shared__ __align__(16) int rows_s[kAspChunk];
    __shared__ __align__(16) float xs_s[B * kAspChunk];
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

    float4 acc[B];
#pragma unroll
    for (int b = 0; b < B; b++) acc[b] = make_float4(0.f, 0.f, 0.f, 0.f);
    pdl_wait();
    for (int r0 = r_begin; r0 < r_end; r0 += kAspChunk) {
        // ---- compaction of x[r0 .. r0+chunk): warp w takes a contiguous quarter ----------------
        constexpr int kSpan = kAspChunk / (kAspThreads / 32);   // 256 rows per warp
        constexpr int kSteps = kSpan / 32;                      // 8
        float xr[B][kSteps]; unsigned bal[kSteps];
        int cnt = 0;
#pragma unroll
        for (int k = 0; k < kSteps; k++) {
            const int row = r0 + warp * kSpan + k * 32 + lane;
            bool any = false;
#pragma unroll
            for (int b = 0; b < B; b++) {
                xr[b][k] = row < r_end ? __ldg(x + b * ldx + row) : 0.0f;
                any = any || xr[b][k] != 0.0f;
            }
            bal[k] = __ballot_sync(kFull, any);
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
            if (bal[k] & (1u << lane)) {
                const int pos = base + __popc(bal[k] & lt);
                rows_s[pos] = r0 + warp * kSpan + k * 32 + lane;
#pragma unroll
                for (int b = 0; b < B; b++) xs_s[b * kAspChunk + pos] = xr[b][k];
            }
            base += __popc(bal[k]);
        }
        __syncthreads();

        // ---- stream the active rows ------------------------------------------------------------
        if (B == 1 && R > 0 && total >= kAspRegsMin) {   // (block-uniform choice: depends on x only through `total`)
            if (col_ok) {
                constexpr int D = R > 0 ? R : 4;
                float4 a[D];
#pragma unroll
                for (int k = 0; k < D; k += 4) {
                    const int4 rw = *reinterpret_cast<const int4 *>(rows_s + k);   // (entries past `total` are stale: not used)
                    a[k + 0] = asp_row_load(Ac + (long long)rw.x * ld, k + 0 < total);
                    a[k + 1] = asp_row_load(Ac + (long long)rw.y * ld, k + 1 < total);
                    a[k + 2] = asp_row_load(Ac + (long long)rw.z * ld, k + 2 < total);
                    a[k + 3] = asp_row_load(Ac + (long long)rw.w * ld, k + 3 < total);
                }
                for (int i0 = 0; i0 < total; i0 += D) {
#pragma unroll
                    for (int k = 0; k < D; k += 4) {
                        const int i = i0 + k;
                        if (i >= total) break;                              // (block-uniform)
                        const float4 xv = *reinterpret_cast<const float4 *>(xs_s + i);   // slots past `total`: their rows are zeros
                        const int j = i + D;
                        int4 rw = make_int4(0, 0, 0, 0);
                        if (j < total) rw = *reinterpret_cast<const int4 *>(rows_s + (j & (kAspChunk - 1)));
                        const float xk[4] = {xv.x, xv.y, xv.z, xv.w};
                        const int rk[4] = {rw.x, rw.y, rw.z, rw.w};
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            if (i + q < total) {
                                acc[0].x = fmaf(a[k + q].x, xk[q], acc[0].x); acc[0].y = fmaf(a[k + q].y, xk[q], acc[0].y);
                                acc[0].z = fmaf(a[k + q].z, xk[q], acc[0].z); acc[0].w = fmaf(a[k + q].w, xk[q], acc[0].w);
                            }
                            a[k + q] = asp_row_load(Ac + (long long)rk[q] * ld, j + q < total);
                        }
                    }
                }
            }
        } else if (col_ok) {
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
#pragma unroll
                for (int b = 0; b < B; b++) {
                    const float xv = xs_s[b * kAspChunk + i];
                    acc[b].x = fmaf(a.x, xv, acc[b].x); acc[b].y = fmaf(a.y, xv, acc[b].y);
                    acc[b].z = fmaf(a.z, xv, acc[b].z); acc[b].w = fmaf(a.w, xv, acc[b].w);
                }
                issue(i + kAspStages);
            }
            cp_async_wait<0>();
        }
    }

    if (splits == 1) {
#pragma unroll
        for (int b = 0; b < B; b++)
            if (col_ok) y_store4(yd, ((size_t)b * ldy + c0) >> 2, acc[b]);
        return;
    }
    const size_t npad = (size_t)gridDim.x * kAspTile;
    const int n_valid = min(kAspTile, N - tile * kAspTile);
#pragma unroll
    for (int b = 0; b < B; b++)
        *reinterpret_cast<float4 *>(partial + ((size_t)b * splits + split) * npad + (size_t)tile * kAspTile + tid * 4) = acc[b];
    __syncthreads();                                      // the row list is dead: reuse it as scratch
    for (int b = 0; b < B; b++) {
        YDst yb = yd;
        for (int k = 0; k < yb.n; k++) yb.p[k] += (size_t)b * ldy;
        if (yb.mc) yb.mc += (size_t)b * ldy;
        split_reduce_finish(yb, partial + (size_t)b * splits * npad, tickets + (size_t)b * gridDim.x, tile, splits, kAspTile,
                            n_valid, npad, &last_flag, reinterpret_cast<float4 *>(rows_s));
        __syncthreads();
    }
}

} // namespace

template <int B, int R = 0>
static int launch_asp_b(spmv_plan *p, const float *d_x, const YDst &yd, cudaStream_t st, long long ldx, long long ldy)
{
    const int smem = (kAspThreads / 32) * kAspStages * 32 * (int)sizeof(float4);
    static int smem_set[16] = {0};   // static + dynamic shared memory exceeds 48 KB for B = 4: always opt in
    if (p->device >= 0 && p->device < 16 && smem_set[p->device] < smem) {
        SPMV_CUDA(cudaFuncSetAttribute(asp_kernel<B, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        smem_set[p->device] = smem;
    }
    SPMV_CUDA(launch_k(asp_kernel<B, R>, p->grid, dim3(kAspThreads), smem, st, p->asp.A, (long long)p->asp.ld, d_x, yd, p->partial,
                       p->tickets, (int)p->M, (int)p->N, p->asp.rows_per_split, p->row_splits, ldx, ldy));
    return SPMV_OK;
}

int launch_asp(spmv_plan *p, const float *d_x, const YDst &yd, cudaStream_t st)
{
    if (p->N == 0) return SPMV_OK;
    // a CTA's row range must be able to hold a long list at all (x decides at run time, chunk by chunk)
    if (kAspRegs > 0 && p->asp.rows_per_split >= 2 * kAspRegsMin) return launch_asp_b<1, kAspRegs>(p, d_x, yd, st, 0, 0);
    return launch_asp_b<1>(p, d_x, yd, st, 0, 0);
}

// Batched form: B in {2, 4} vectors in one pass over A's rows.
int launch_asp_batch(spmv_plan *p, const float *d_x, long long ldx, const YDst &yd, long long ldy, int B, cudaStream_t st)
{
    if (p->N == 0) return SPMV_OK;
    if (B == 2) return launch_asp_b<2>(p, d_x, yd, st, ldx, ldy);
    if (B == 4) return launch_asp_b<4>(p, d_x, yd, st, ldx, ldy);
    return SPMV_ERR_UNSUPPORTED;
}

// Round 2 tried two re-decompositions, both measured slower or equal on a B200 and removed again
// (profiles/r02_notes.md): (a) a flat (tile, row) sequence cut into 2-5 CTAs per SM with 16/32-deep
// rings (config 2: 25.9-29.5 us against 23.7 here; config 0: 11.8-16.5 against 11.7) — more bytes in
// flight through cp.async do not help; (b) one warp per (128-column subtile, row range), 8 rows in
// flight as plain register loads, 32 warps per SM, a separate reduce kernel (config 2: 23.5 us;
// config 0 / 3: 13.9 / 14.1 against 11.7 / 9.4): with ~50 active rows per warp the per-warp ramp and
// the extra launch outweigh the cheaper loads.  A plain read of the same pattern reaches 18 us on
// config 2 (tools/ubench/rowpattern.cu); the remaining gap is the x-dependent head and the tail.
// grid = (ceil(N/512), row splits): about 1.6 CTAs per SM, all resident at once — every CTA pays a
// fixed few microseconds (x compaction, first rows, split reduction), so more splits are slower —
// with the rows per split rounded DOWN to a multiple of 32 (at least 64).  Measured with dependent
// launch on (profiles/r01_notes.md): config 2 8 splits 23.8 us, 12 splits 24.3, 16 splits 26.9;
// config 3 32 splits 9.4 us, 41 splits 10.3; config 0 32 splits 11.7 us, 16 splits 12.2.
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
    else splits = std::max(1, (16 * p->sm_count / 10 + p->col_tiles / 2) / std::max(1, p->col_tiles));
    int rps;
    if ((o && o->row_splits > 0) || splits == 1) rps = ((int)((M + splits - 1) / splits) + 31) / 32 * 32;
    else rps = std::max(64, (int)(M / splits) / 32 * 32);
    splits = (int)((M + rps - 1) / rps);
    p->asp.rows_per_split = rps;
    p->row_splits = splits;
    p->grid = dim3((unsigned)std::max(1, p->col_tiles), (unsigned)splits, 1);
    return alloc_split_scratch(p, kAspMaxBatch);
}

} // namespace spmv
