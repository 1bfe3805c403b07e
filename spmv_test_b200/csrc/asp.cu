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
#include <cstdlib>
#include <cstring>

#include <cuda.h>      // CUtensorMap and its enums only: the encoder is fetched through cudaGetDriverEntryPoint (no libcuda link)

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

// ---- TMA row gather (cp.async.bulk.tensor.2d ... tile::gather4) ---------------------------------------------
// The single-vector kernel with the rows moved by the TMA engine instead of the SM's load/store units: a fifth warp's
// lane 0 walks the compacted row list four rows at a time and issues, per group, two gather4 copies (4 arbitrary rows
// x 256 columns = 4 KB each; the tensor map's box is 256 x 1) into a ring of kAspTmaStages 8 KB stages guarded by
// full (transaction bytes) / empty (128 arrivals) mbarriers; the four consumer warps read their 16 bytes of each row
// back and do the same FMAs in the same order as asp_kernel, so y is bit-identical to it.  Micro-benchmark of the bare
// access pattern on config 2 (tools/ubench/gather4.cu, same box): 18.75 us against 20.77 us for 32 rows in flight in
// registers (4 stages 20.3, 12 stages 19.3).  In the real kernel — compaction in front, FMAs, split reduction behind
// — it LOSES to the register / ring kernel: config 2 / 0 / 3 24.22 / 12.14 / 9.26 us against 21.91 / 11.55 / 9.21
// (same box, parity tests green on both), so it is not the default: SPMV_ASP_TMA=1 selects it.
#ifndef SPMV_ASP_TMA_DEFAULT
#define SPMV_ASP_TMA_DEFAULT 0
#endif
#ifndef SPMV_ASP_TMA_STAGES
#define SPMV_ASP_TMA_STAGES 8
#endif
constexpr int kAspTmaStages = SPMV_ASP_TMA_STAGES;
constexpr int kAspTmaThreads = kAspThreads + 32;
struct alignas(64) AspTmap { unsigned char b[128]; };

__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(kAspTmaThreads)
asp_tma_kernel(const __grid_constant__ AspTmap tm, const float *__restrict__ x, const YDst yd, float *__restrict__ partial,
               unsigned *__restrict__ tickets, int M, int N, int rows_per_split, int splits)
{
    static_assert(kAspThreads == 128 && kAspTile == 512, "two 256-column boxes per tile, four consumer warps");
    __shared__ __align__(16) int rows_s[kAspChunk];
    __shared__ __align__(16) float xs_s[kAspChunk];
    __shared__ int wcnt[kAspThreads / 32];
    __shared__ int last_flag;
    __shared__ __align__(8) uint64_t full[kAspTmaStages], empty[kAspTmaStages];
    extern __shared__ unsigned char tma_raw[];
    unsigned char *stage = tma_raw + ((1024u - (smem_u32(tma_raw) & 1023u)) & 1023u);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x, split = blockIdx.y;
    const int c0 = tile * kAspTile + tid * 4;
    const bool consumer = warp < kAspThreads / 32;
    const bool col_ok = consumer && c0 < N;
    const int halves = tile * kAspTile + 256 < N ? 2 : 1;  // a tile's second box may lie past N: not requested
    const int r_begin = split * rows_per_split;
    const int r_end = min(M, r_begin + rows_per_split);
    if (tid == 0) {
        for (int s = 0; s < kAspTmaStages; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], kAspThreads); }
        fence_mbar_init();
    }
    __syncthreads();

    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    pdl_wait();
    unsigned gdone = 0;                                   // four-row groups through the ring so far (same in every thread)
    for (int r0 = r_begin; r0 < r_end; r0 += kAspChunk) {
        constexpr int kSpan = kAspChunk / (kAspThreads / 32);
        constexpr int kSteps = kSpan / 32;
        float xr[kSteps]; unsigned bal[kSteps];
        int cnt = 0;
        if (consumer) {
#pragma unroll
            for (int k = 0; k < kSteps; k++) {
                const int row = r0 + warp * kSpan + k * 32 + lane;
                xr[k] = row < r_end ? __ldg(x + row) : 0.0f;
                bal[k] = __ballot_sync(kFull, xr[k] != 0.0f);
                cnt += __popc(bal[k]);
            }
        }
        __syncthreads();                                  // previous chunk's list fully consumed (by the producer too)
        if (consumer && lane == 0) wcnt[warp] = cnt;
        __syncthreads();
        int base = 0, total = 0;
#pragma unroll
        for (int w = 0; w < kAspThreads / 32; w++) {
            const int cw = wcnt[w];
            if (w < warp) base += cw;
            total += cw;
        }
        if (consumer) {
            const unsigned lt = (1u << lane) - 1u;
#pragma unroll
            for (int k = 0; k < kSteps; k++) {
                if (bal[k] & (1u << lane)) {
                    const int pos = base + __popc(bal[k] & lt);
                    rows_s[pos] = r0 + warp * kSpan + k * 32 + lane;
                    xs_s[pos] = xr[k];
                }
                base += __popc(bal[k]);
            }
        }
        __syncthreads();

        const int ngroups = (total + 3) >> 2;
        if (!consumer) {
            if (lane == 0) {
                for (int g = 0; g < ngroups; g++) {
                    const unsigned gg = gdone + (unsigned)g;
                    const int s = (int)(gg % kAspTmaStages);
                    if (gg >= (unsigned)kAspTmaStages) mbar_wait(&empty[s], ((gg / kAspTmaStages) - 1u) & 1u);
                    mbar_expect_tx(&full[s], (uint32_t)halves * 4096u);
                    int rr[4];
#pragma unroll
                    for (int q = 0; q < 4; q++) rr[q] = rows_s[min(4 * g + q, total - 1)];   // a short last group repeats its last row
                    for (int h = 0; h < halves; h++)
                        asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes "
                                     "[%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                                     ::"r"(smem_u32(stage + s * 8192 + h * 4096)), "l"(&tm), "r"(tile * kAspTile + h * 256), "r"(rr[0]),
                                     "r"(rr[1]), "r"(rr[2]), "r"(rr[3]), "r"(smem_u32(&full[s])) : "memory");
                }
            }
        } else {
            for (int g = 0; g < ngroups; g++) {
                const unsigned gg = gdone + (unsigned)g;
                const int s = (int)(gg % kAspTmaStages);
                mbar_wait(&full[s], (gg / kAspTmaStages) & 1u);
                if (col_ok) {
                    const float4 *rowp = reinterpret_cast<const float4 *>(stage + s * 8192 + (tid >> 6) * 4096) + (tid & 63);
                    const float4 xv = *reinterpret_cast<const float4 *>(xs_s + 4 * g);     // (slots past `total` are stale: not used)
                    const float xk[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        if (4 * g + q < total) {
                            const float4 a = rowp[q * 64];
                            acc.x = fmaf(a.x, xk[q], acc.x); acc.y = fmaf(a.y, xk[q], acc.y);
                            acc.z = fmaf(a.z, xk[q], acc.z); acc.w = fmaf(a.w, xk[q], acc.w);
                        }
                    }
                }
                mbar_arrive(&empty[s]);
            }
        }
        gdone += (unsigned)ngroups;
    }

    if (splits == 1) {
        if (col_ok) y_store4(yd, (size_t)c0 >> 2, acc);
        return;
    }
    const size_t npad = (size_t)gridDim.x * kAspTile;
    const int n_valid = min(kAspTile, N - tile * kAspTile);
    if (consumer) *reinterpret_cast<float4 *>(partial + (size_t)split * npad + (size_t)tile * kAspTile + tid * 4) = acc;
    __syncthreads();                                      // the row list is dead: reuse it as scratch
    split_reduce_finish(yd, partial, tickets, tile, splits, kAspTile, n_valid, npad, &last_flag, reinterpret_cast<float4 *>(rows_s));
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

// The tensor map of this plan's A for asp_tma_kernel: box = 256 columns x 1 row (what tile::gather4 requires: a box of
// four rows is an illegal instruction), no swizzle, 128-byte L2 promotion.  Encoded once per (plan, buffer).
static bool asp_tma_ready(spmv_plan *p)
{
    DevAsp &a = p->asp;
    if (a.tma_state == 1 && a.tmap_for == a.A) return true;
    if (a.tma_state < 0 && a.tmap_for == a.A) return false;
    a.tmap_for = a.A;
    a.tma_state = -1;
    static_assert(sizeof(CUtensorMap) == sizeof(a.tmap), "tensor maps are 128 bytes");
    typedef CUresult (*EncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeTiled encode = [] {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) fn = nullptr;
        cudaGetLastError();
        return reinterpret_cast<EncodeTiled>(fn);
    }();
    if (!encode || p->N < 256 || p->M < 1 || (a.ld & 3) || (reinterpret_cast<uintptr_t>(a.A) & 15)) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)p->N, (cuuint64_t)p->M};
    const cuuint64_t strides[1] = {(cuuint64_t)a.ld * 4};
    const cuuint32_t box[2] = {256u, 1u};
    const cuuint32_t estr[2] = {1u, 1u};
    if (encode(reinterpret_cast<CUtensorMap *>(a.tmap), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, a.A, dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return false;
    a.tma_state = 1;
    return true;
}

// 1: single-vector asp calls go through asp_tma_kernel when the plan's matrix can be described by a tensor map
static bool asp_tma_wanted()
{
    static const int on = [] { const char *e = std::getenv("SPMV_ASP_TMA"); return e ? std::atoi(e) : SPMV_ASP_TMA_DEFAULT; }();
    return on != 0;
}

static int launch_asp_tma(spmv_plan *p, const float *d_x, const YDst &yd, cudaStream_t st)
{
    const int smem = kAspTmaStages * 8192 + 1024;
    static int smem_set[16] = {0};
    const bool cached = p->device >= 0 && p->device < 16;
    if (!cached || !smem_set[p->device]) {                // (beyond 16 devices: set on every call, the opt-in is mandatory)
        SPMV_CUDA(cudaFuncSetAttribute(asp_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        if (cached) smem_set[p->device] = 1;
    }
    AspTmap tm;
    std::memcpy(tm.b, p->asp.tmap, sizeof tm.b);
    SPMV_CUDA(launch_k(asp_tma_kernel, p->grid, dim3(kAspTmaThreads), (size_t)smem, st, tm, d_x, yd, p->partial, p->tickets, (int)p->M,
                       (int)p->N, p->asp.rows_per_split, p->row_splits));
    return SPMV_OK;
}

int launch_asp(spmv_plan *p, const float *d_x, const YDst &yd, cudaStream_t st)
{
    if (p->N == 0) return SPMV_OK;
    if (asp_tma_wanted() && asp_tma_ready(p)) return launch_asp_tma(p, d_x, yd, st);
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
