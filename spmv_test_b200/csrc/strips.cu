// strips.cu — activation+weight-sparse SGEMV on the row-strip format (formats.hpp: HostStrips).
//
// The awsp variant (reference awsp_kernel_v0/v1/v2, awsp.cu:5-317; awsp_ref_kernel, awsp_ref.cu:6-185)
// for very sparse matrices such as BASELINE config 5 (1 % dense, half of x non-zero), where a row
// segment of a 2048-column slab holds ~20 non-zeros.  The two earlier forms each give up one of the
// variant's promises there: the row-addressable multi-row schedule (panel.cu) is instruction bound
// (a 128-entry chunk mixes ~6 rows that may share columns, so it retires in ~7 passes), the
// lane-owned blocks read every stored non-zero whatever x is.  Here
//   * a 16-warp CTA owns (band of 16 strips, row range); warp w owns strip w of every row, so the
//     16 warps together read the contiguous 2-3 KB band segment of an active row (DRAM-friendly
//     pieces) while each warp only ever touches its private strip accumulators in shared memory;
//   * the CTA compacts its x slice once (ballot + popc, order preserving: the x != 0.0f test of
//     awsp.cu:98,127) into a shared list of active rows; rows with x == 0 are never addressed;
//   * per active row a warp loads its strip segment — one 8-byte entry per lane, idle lanes get
//     the all-zero entry — straight into registers, 32 rows in flight per warp (a cp.async ring is
//     kept as the SPMV_STRIP_REGS=0 build: it costs LSU wavefronts twice and was 45 % slower), and
//     retires it in ONE pass: the entries of one row are distinct columns, so no two lanes meet
//     in an accumulator (no passes, no votes, no predicates: an idle lane adds 0 to accumulator 0,
//     columns are stored + 1);
//   * per-row scalars (segment start, length, x) are fetched lane = row, two to three 32-row
//     batches ahead, into a small per-warp table, so the row loop reads them with broadcast loads;
//   * the row ranges of a band are added in range order by strips_reduce_kernel (dependent launch).
// Deterministic: no atomics at all.
#include <algorithm>

#include "common.cuh"
#include "plan.hpp"

namespace spmv {

namespace {

constexpr int kStripWarps = kStripsPerBand;               // warp w <-> strip w of the band
constexpr int kStripThreads = kStripWarps * 32;
#ifndef SPMV_STRIP_STAGES
#define SPMV_STRIP_STAGES 32
#endif
constexpr int kStripStages = SPMV_STRIP_STAGES;           // rows in flight per warp (divides 32)
// Where the rows in flight wait: 1 = in registers (four groups of four 8-byte loads per lane), 0 = in
// a per-warp cp.async ring in shared memory.  The kernel is bound by shared-memory wavefronts (the
// accumulator read-modify-writes); a ring adds 4 wavefronts per row for the copy-in and 2 for the
// read-back to the 6 of the updates themselves (ncu, profiles/r02_notes.md).
#ifndef SPMV_STRIP_REGS
#define SPMV_STRIP_REGS 1
#endif
constexpr bool kStripRegs = SPMV_STRIP_REGS != 0;
#ifndef SPMV_STRIP_PREFETCH
#define SPMV_STRIP_PREFETCH 0
#endif
constexpr bool kStripPrefetch = SPMV_STRIP_PREFETCH != 0;
// batches of per-row scalars (segment offsets, x) in flight ahead of the one being parked: with one, the park at the
// top of a batch was the kernel's largest single stall site (ncu: 1125 of ~5400 samples on the subtraction that first
// uses the loaded offsets) — under a saturated LSU pipe a global load needs more than one batch time to return.
// Two batches ahead removes that stall and changes the time by 0.1 % (93.87 against 93.97-94.07 us per slab, three
// alternating runs on one box): the pipe's wavefronts, not the warps' stalls, set the time.
#ifndef SPMV_STRIP_META_AHEAD
#define SPMV_STRIP_META_AHEAD 2
#endif
constexpr int kStripMetaAhead = SPMV_STRIP_META_AHEAD;
// (four rows share one cp.async commit group and one set of broadcast loads of their scalars)
constexpr int kStripSub = 2048;                           // rows compacted per pass
constexpr int kStripSpan = kStripSub / kStripWarps;       // rows a warp compacts: 128
constexpr int kStripSteps = kStripSpan / 32;

__host__ __device__ constexpr int strip_warp_bytes(int sw) { return (sw + 32) * 4 + (kStripRegs ? 0 : kStripStages * 32 * 8); }
__host__ __device__ constexpr int strip_smem_bytes(int sw) { return kStripWarps * (strip_warp_bytes(sw) + 128 * 12) + kStripSub * (2 + 4); }

// The row loop's ring traffic as volatile asm WITHOUT a memory clobber: volatile statements keep
// their program order among themselves (ring read -> refill of the same slot -> commit -> wait ->
// next ring read), while the compiler stays free to schedule the accumulator and row-table
// accesses around them (with the clobber it drains every outstanding shared-memory load before
// each copy: three extra issue slots per row).
__device__ __forceinline__ void ring_copy8(uint32_t smem_dst, uint64_t gmem_src, bool on)   // off: the slot is zero-filled
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(smem_dst), "l"(gmem_src), "r"(on ? 8 : 0));
}
__device__ __forceinline__ uint2 ring_read8(uint32_t smem_src)
{
    uint2 r;
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "r"(smem_src));
    return r;
}
// register staging: one 8-byte entry per lane, idle lanes get the all-zero entry; volatile so the
// load is issued where it is written (16 rows before its use), not sunk to the use
__device__ __forceinline__ uint2 entry_load(uint64_t gmem_src, bool on)
{
    uint2 r;
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %3, 0;\n\tmov.b32 %0, 0;\n\tmov.b32 %1, 0;\n\t"
                 "@p ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];\n\t}"
                 : "=&r"(r.x), "=&r"(r.y) : "l"(gmem_src), "r"((int)on));
    return r;
}
__device__ __forceinline__ void ring_commit() { asm volatile("cp.async.commit_group;"); }
template <int N> __device__ __forceinline__ void ring_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

__global__ void __launch_bounds__(kStripThreads, 1)
strips_kernel(const uint2 *__restrict__ ent, const uint32_t *__restrict__ soff, const float *__restrict__ x,
              const YDst yd, float *__restrict__ partial, int M, int N, int sw, int R)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int wcnt[kStripWarps];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int band = blockIdx.x / R, part = blockIdx.x - band * R;
    const int row_a = (int)((long long)part * M / R), row_b = (int)((long long)(part + 1) * M / R);
    unsigned char *wbase = smem_raw + (size_t)warp * strip_warp_bytes(sw);
    float *acc = reinterpret_cast<float *>(wbase);        // [0]: idle lanes; [1 .. sw]: the strip's columns
    uint2 *ring = reinterpret_cast<uint2 *>(wbase + (size_t)(sw + 32) * 4);
    uint16_t *rows_s = reinterpret_cast<uint16_t *>(smem_raw + (size_t)kStripWarps * strip_warp_bytes(sw));
    float *xs_s = reinterpret_cast<float *>(rows_s + kStripSub);
    uint2 *tab_sl = reinterpret_cast<uint2 *>(xs_s + kStripSub) + warp * 128;          // (start, length) of four 32-row batches
    float *tab_x = reinterpret_cast<float *>(reinterpret_cast<uint2 *>(xs_s + kStripSub) + kStripWarps * 128) + warp * 128;   // their x
    const unsigned lt = (1u << lane) - 1u;

    for (int c = lane; c < sw + 32; c += 32) acc[c] = 0.0f;
    const uint32_t *so = soff + ((size_t)band * (kStripsPerBand + 1) + warp) * M;   // strip-major: this strip's starts, row by row
    // The reduce kernel of the previous call releases its dependents at entry, so this CTA is usually resident
    // and zeroed before that call has finished: the offset records of its first rows (static plan data) are
    // pulled into L2 meanwhile; x and the partial rows may belong to the previous kernels and wait.
    for (int r = lane * 32; r < min(row_b - row_a, kStripSub) + 32; r += 32 * 32) {
        const int row = min(row_a + r, M - 1);
        asm volatile("prefetch.global.L2 [%0];" ::"l"(so + row));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(so + M + row));
    }
    pdl_wait();
    uint64_t ent_lane = reinterpret_cast<uint64_t>(ent + lane);   // opaque to the optimiser: one IMAD.WIDE per address
    asm volatile("" : "+l"(ent_lane));
    const uint32_t ring_lane = smem_u32(ring + lane);

    for (int sub0 = row_a; sub0 < row_b; sub0 += kStripSub) {
        // ---- activation compaction of rows [sub0, sub0 + kStripSub): warp w takes 128 of them ------
        float xr[kStripSteps]; unsigned bal[kStripSteps];
        int cnt = 0;
#pragma unroll
        for (int k = 0; k < kStripSteps; k++) {
            const int row = sub0 + warp * kStripSpan + k * 32 + lane;
            xr[k] = row < row_b ? __ldg(x + row) : 0.0f;
            bal[k] = __ballot_sync(kFull, xr[k] != 0.0f);
            cnt += __popc(bal[k]);
        }
        __syncthreads();                                  // every warp has finished the previous list
        if (lane == 0) wcnt[warp] = cnt;
        __syncthreads();
        int base = 0, total = 0;
#pragma unroll
        for (int w = 0; w < kStripWarps; w++) {
            const int cw = wcnt[w];
            if (w < warp) base += cw;
            total += cw;
        }
#pragma unroll
        for (int k = 0; k < kStripSteps; k++) {
            if (bal[k] & (1u << lane)) {
                const int pos = base + __popc(bal[k] & lt);
                rows_s[pos] = (uint16_t)(warp * kStripSpan + k * 32 + lane);
                xs_s[pos] = xr[k];
            }
            base += __popc(bal[k]);
        }
        __syncthreads();

        // ---- this warp's strip of every active row ------------------------------------------------
        // Per-row scalars (segment start, length, x) are fetched lane = row, 32 rows at a time and
        // two batches ahead of their use, and parked in a small per-warp table in shared memory
        // (structure of arrays), so a group of four rows reads them with broadcast 128-bit loads.
        // A group's loads (scalars, ring slots) are all issued before its first accumulator
        // update: the updates themselves must stay in row order (two rows may share a column).
        struct Meta { uint32_t st, end, x; };
        auto load_meta = [&](int b0) {                    // lane i: row b0 + i of the list; raw loads, used a batch later
            Meta m{0u, 0u, 0u};
            const int i = b0 + lane;
            if (i < total) {
                const uint32_t *p = so + (sub0 + rows_s[i]);          // 32 nearby rows: a few sectors per request
                m.st = __ldg(p);
                m.end = __ldg(p + M);                                // the next strip's start (strip 16: the end of the row)
                m.x = __float_as_uint(xs_s[i]);
            }
            return m;
        };
        auto park = [&](uint2 *sl, float *xs, const Meta &m) {
            sl[lane] = make_uint2(m.st, m.end - m.st);
            xs[lane] = __uint_as_float(m.x);
        };
        struct Four { uint4 a, b; };                      // (start, length) of four rows
        auto group_meta = [&](const uint2 *sl, int u) {
            return Four{*reinterpret_cast<const uint4 *>(sl + u), *reinterpret_cast<const uint4 *>(sl + u + 2)};
        };
        auto issue_group = [&](const Four &f, int slot) {                       // four rows -> slots slot .. slot+3
            const uint4 a = f.a, b = f.b;
            ring_copy8(ring_lane + (slot + 0) * 256, ent_lane + (uint64_t)a.x * 8u, lane < (int)a.y);
            ring_copy8(ring_lane + (slot + 1) * 256, ent_lane + (uint64_t)a.z * 8u, lane < (int)a.w);
            ring_copy8(ring_lane + (slot + 2) * 256, ent_lane + (uint64_t)b.x * 8u, lane < (int)b.y);
            ring_copy8(ring_lane + (slot + 3) * 256, ent_lane + (uint64_t)b.z * 8u, lane < (int)b.w);
            ring_commit();
        };
        auto load_group = [&](const Four &f, uint2 (&e)[4]) {                   // register staging
            e[0] = entry_load(ent_lane + (uint64_t)f.a.x * 8u, lane < (int)f.a.y);
            e[1] = entry_load(ent_lane + (uint64_t)f.a.z * 8u, lane < (int)f.a.w);
            e[2] = entry_load(ent_lane + (uint64_t)f.b.x * 8u, lane < (int)f.b.y);
            e[3] = entry_load(ent_lane + (uint64_t)f.b.z * 8u, lane < (int)f.b.w);
        };
        auto update_group = [&](const float *xs, int u, const uint2 (&e)[4]) {
            const float4 xv = *reinterpret_cast<const float4 *>(xs + u);
            const float xk[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
            for (int k = 0; k < 4; k++) {
                // idle lanes hold the all-zero entry: value 0 into accumulator 0, which no real entry uses (columns are stored + 1)
                float *a = acc + e[k].y;
                *a = fmaf(__uint_as_float(e[k].x), xk[k], *a);
                __syncwarp();                             // the next row may hit the same columns from other lanes
            }
        };
        auto retire_group = [&](const float *xs, int u, int slot) {
            const float4 xv = *reinterpret_cast<const float4 *>(xs + u);
            uint2 e[4];
#pragma unroll
            for (int k = 0; k < 4; k++) e[k] = ring_read8(ring_lane + (slot + k) * 256);   // a lane reads back what it copied itself
            const float xk[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
            for (int k = 0; k < 4; k++) {
                // idle lanes were zero-filled: value 0 into accumulator 0, which no real entry uses (columns are stored + 1)
                float *a = acc + e[k].y;
                *a = fmaf(__uint_as_float(e[k].x), xk[k], *a);
                __syncwarp();                             // the next row may hit the same columns from other lanes
            }
        };
        // the table is a ring of four 32-row batches: batch i sits in slot i & 3; while batch i is being
        // retired, batch i+1 is being loaded from and batch i+2 prefetched from
        auto sl_of = [&](int bi) { return tab_sl + (bi & 3) * 32; };
        auto xs_of = [&](int bi) { return tab_x + (bi & 3) * 32; };
        // L2 prefetch of the rows 32 beyond the ones being loaded: one lane per 32-byte sector of the
        // segment (segments start on sector boundaries).  No data returns to the SM, so it costs an
        // issue slot but none of the data-pipe wavefronts the kernel is bound by, and the demand
        // loads that follow find their sectors in L2.
        auto prefetch_group = [&](const Four &f) {
            if (!kStripPrefetch || (lane & 3)) return;
            if (lane < (int)f.a.y) asm volatile("prefetch.global.L2 [%0];" ::"l"(ent_lane + (uint64_t)f.a.x * 8u));
            if (lane < (int)f.a.w) asm volatile("prefetch.global.L2 [%0];" ::"l"(ent_lane + (uint64_t)f.a.z * 8u));
            if (lane < (int)f.b.y) asm volatile("prefetch.global.L2 [%0];" ::"l"(ent_lane + (uint64_t)f.b.x * 8u));
            if (lane < (int)f.b.w) asm volatile("prefetch.global.L2 [%0];" ::"l"(ent_lane + (uint64_t)f.b.z * 8u));
        };
        park(sl_of(0), xs_of(0), load_meta(0));
        park(sl_of(1), xs_of(1), load_meta(32));
        Meta ahead = load_meta(64);
        Meta ahead2 = kStripMetaAhead > 1 ? load_meta(96) : Meta{0u, 0u, 0u};   // (a second batch of scalars in flight)
        __syncwarp();
        constexpr int kGroups = kStripStages / 4;         // groups of four rows in flight
        uint2 E[kGroups][4];
#pragma unroll
        for (int u = 0; u < kStripStages; u += 4) {
            const Four f = u < 32 ? group_meta(sl_of(0), u) : group_meta(sl_of(1), u - 32);
            if (kStripRegs) load_group(f, E[u / 4]);
            else issue_group(f, u);
        }
#pragma unroll 1
        for (int bi = 0; bi * 32 < total; bi++) {
            park(sl_of(bi + 2), xs_of(bi + 2), ahead);    // (its loads had one or two whole batches to land)
            if (kStripMetaAhead > 1) { ahead = ahead2; ahead2 = load_meta((bi + 4) * 32); }
            else ahead = load_meta((bi + 3) * 32);
            __syncwarp();
            const uint2 *sl0 = sl_of(bi), *sl1 = sl_of(bi + 1), *sl2 = sl_of(bi + 2);
            const float *xs0 = xs_of(bi);
#pragma unroll
            for (int u = 0; u < 32; u += 4) {
                // (the refill's scalars are read before the accumulator updates, not behind them)
                constexpr int S = kStripStages;
                const Four nx = u + S < 32 ? group_meta(sl0, u + S) : u + S < 64 ? group_meta(sl1, u + S - 32) : group_meta(sl2, u + S - 64);
                if (kStripPrefetch) prefetch_group(u + S + 32 < 64 ? group_meta(sl1, u + S) : group_meta(sl2, u + S - 32));
                if (kStripRegs) {
                    update_group(xs0, u, E[(u / 4) % kGroups]);
                    load_group(nx, E[(u / 4) % kGroups]);
                } else {
                    ring_wait<kStripStages / 4 - 1>();    // the oldest group of rows has landed
                    retire_group(xs0, u, u % kStripStages);
                    issue_group(nx, u % kStripStages);
                }
            }
            // segments longer than a window (under 1 % of them by the strip-width rule): the entries past
            // the first 32, straight from global memory, after the batch (a fixed order all the same)
            const uint2 mine = sl0[lane];
            const float myx = xs0[lane];
            unsigned long_rows = __ballot_sync(kFull, mine.y > 32u);
            while (long_rows) {
                const int u = __ffs(long_rows) - 1;
                long_rows &= long_rows - 1;
                const uint32_t st = __shfl_sync(kFull, mine.x, u), cn = __shfl_sync(kFull, mine.y, u);
                const float xv = __shfl_sync(kFull, myx, u);
                for (uint32_t o = 32 + lane; o < cn; o += 32) {
                    const uint2 f = __ldg(ent + st + o);
                    acc[f.y] = fmaf(__uint_as_float(f.x), xv, acc[f.y]);
                }
                __syncwarp();
            }
        }
        if (!kStripRegs) ring_wait<0>();                  // (only empty groups are left)
    }

    // ---- this warp's strip of the band: y, or one partial row per (band, row range) --------------
    __syncwarp();
    const size_t band_cols = (size_t)sw * kStripsPerBand;
    const size_t col0 = (size_t)band * band_cols + (size_t)warp * sw;
    const float *ac = acc + 1;                            // accumulator 0 belongs to the idle lanes
    if (R == 1) {
        for (int c = lane * 4; c < sw; c += 128)
            if (col0 + c < (size_t)N) y_store4(yd, (col0 + c) >> 2, make_float4(ac[c], ac[c + 1], ac[c + 2], ac[c + 3]));
    } else {
        float *dst = partial + (size_t)blockIdx.x * band_cols + (size_t)warp * sw;
        for (int c = lane * 4; c < sw; c += 128) *reinterpret_cast<float4 *>(dst + c) = make_float4(ac[c], ac[c + 1], ac[c + 2], ac[c + 3]);
    }
}

// y[col] = sum over the band's row ranges.  A CTA handles 32 float4 columns with 8 row groups: group k
// adds ranges k, k+8, ... (all of its loads in flight at once for up to 64 ranges), then the 8 group
// sums are added in group order — the order depends only on (shape, R), never on timing.  (One thread
// per column with 8 loads in flight needed five dependent L2 round trips for 37 ranges: ~5 us per call.)
__global__ void __launch_bounds__(256)
strips_reduce_kernel(const float *__restrict__ partial, const YDst yd, int N, int band_cols, int R)
{
    __shared__ float4 sums[256];
    pdl_trigger();                                        // the next kernel's CTAs may come up (and wait) while this one runs
    pdl_wait();
    const int v = threadIdx.x & 31, k = threadIdx.x >> 5;
    const size_t i4 = (size_t)blockIdx.x * 32 + v;
    const size_t col = i4 * 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (col < (size_t)N) {
        const size_t band = col / band_cols, within = col - band * band_cols;
        const float4 *p = reinterpret_cast<const float4 *>(partial + band * R * (size_t)band_cols + within);
        const size_t stride4 = (size_t)band_cols >> 2;
        for (int r = k; r < R; r += 8 * kRedBatch) {
            float4 t[kRedBatch];
#pragma unroll
            for (int u = 0; u < kRedBatch; u++)
                t[u] = r + 8 * u < R ? __ldcg(p + (size_t)(r + 8 * u) * stride4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int u = 0; u < kRedBatch; u++) acc = f4_add(acc, t[u]);
        }
    }
    sums[threadIdx.x] = acc;
    __syncthreads();
    if (k == 0 && col < (size_t)N) {
        float4 s = sums[v];
#pragma unroll
        for (int g = 1; g < 8; g++) s = f4_add(s, sums[g * 32 + v]);
        y_store4(yd, i4, s);
    }
}

} // namespace

int launch_strips(spmv_plan *p, const float *d_x, const YDst &yd, cudaStream_t st)
{
    if (p->N == 0) return SPMV_OK;
    if (p->M == 0) {
        for (int k = 0; k < yd.n; k++) SPMV_CUDA(cudaMemsetAsync(yd.p[k], 0, (size_t)p->N * sizeof(float), st));
        return SPMV_OK;
    }
    static int smem_set[16] = {0};
    if (p->device >= 0 && p->device < 16 && smem_set[p->device] < p->smem) {
        SPMV_CUDA(cudaFuncSetAttribute(strips_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, p->smem));
        smem_set[p->device] = p->smem;
    }
    const DevStrips &d = p->strips;
    SPMV_CUDA(launch_k(strips_kernel, p->grid, dim3(kStripThreads), (size_t)p->smem, st, d.ent, d.soff, d_x, yd, p->partial,
                       (int)p->M, (int)p->N, d.strip_cols, d.ctas_per_band));
    if (d.ctas_per_band > 1) {
        const int band_cols = d.strip_cols * kStripsPerBand;
        const unsigned blocks = (unsigned)((p->N / 4 + 31) / 32);
        SPMV_CUDA(launch_k(strips_reduce_kernel, dim3(blocks), dim3(256), 0, st, (const float *)p->partial, yd, (int)p->N, band_cols,
                           d.ctas_per_band));
    }
    return SPMV_OK;
}

// One CTA per SM (the strip accumulators and rings of 16 warps fill the shared memory), a whole
// number of CTAs per band, every CTA a contiguous row range of its band.
int configure_strips(spmv_plan *p, const HostStrips &h, const spmv_options_t *o)
{
    DevStrips &d = p->strips;
    d.strip_cols = h.strip_cols; d.bands = h.bands;
    const int smem_cap = p->max_smem_optin > 0 ? p->max_smem_optin : 227 * 1024;
    p->smem = strip_smem_bytes(h.strip_cols);
    if (p->smem > smem_cap)
        return set_error(SPMV_ERR_UNSUPPORTED, "strips: %d bytes of shared memory exceed the device limit", p->smem);
    p->block = kStripThreads;
    const int bands = std::max(1, h.bands);
    int64_t R = std::max<int64_t>(1, p->sm_count / bands);
    if (o && o->row_splits > 0) R = o->row_splits;
    R = std::max<int64_t>(1, std::min<int64_t>(R, (std::max<int64_t>(h.M, 1) + 63) / 64));   // at least 64 rows per CTA
    d.ctas_per_band = (int)R;
    p->row_splits = (int)R;
    p->tile_width = h.strip_cols * kStripsPerBand;
    p->col_tiles = bands;
    p->grid = dim3((unsigned)(bands * R), 1, 1);
    p->kernels_per_run = R > 1 ? 2 : 1;
    p->partial = nullptr; p->tickets = nullptr;
    if (R > 1) return alloc_panel_scratch(p, (size_t)bands * R * p->tile_width, 1);
    return SPMV_OK;
}

} // namespace spmv
