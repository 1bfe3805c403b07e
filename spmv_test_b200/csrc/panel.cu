// panel.cu — activation+weight-sparse SGEMV on the row-panel format (formats.hpp: HostPanel).
//
// Replaces awsp_kernel_v0/v1/v2 (reference awsp.cu:5-317), awsp_ref_kernel (awsp_ref.cu:6-185)
// and csr_tiling_kernel (csr_tiling.cu:24-114).  The reference gives every 32-column slab to
// one CTA, walks a per-row bitmap (word -> popc -> address -> one 4-byte load per lane) and
// uses x only as a load predicate: every bitmap word and every x is still read.  Here
//   * the (slab, row) pairs form one flat sequence cut into equal ranges, one per CTA (one
//     resident wave, whole CTAs per slab where possible); a range is one or more pieces
//     (slab, row range);
//   * per piece, lane l of a warp looks at row l of a 32-row block: x[row] and the segment's
//     group range.  ballot(x != 0 && segment non-empty) is the activation compaction: rows with
//     x == 0 are never visited, so their values/indices are never read from HBM; the active
//     rows are dealt to the CTA's warps (whole blocks per warp, or by rank for short pieces);
//   * a visited segment is streamed as 128-bit groups (float4 values + 4 packed column ids),
//     32 groups per chunk, kStages chunks in flight per warp through a cp.async ring in shared
//     memory (commit/wait groups give a true FIFO; a register ring collapses to one load in
//     flight because its loads share scoreboard slots — measured, profiles/r01_notes.md);
//     short segments (config 5) are laid end to end, 32 groups per chunk, and retired row by row;
//   * products are accumulated into a per-warp fp32 accumulator row in shared memory
//     (columns inside one segment are distinct, segments are consumed in ascending row order,
//     warps never share an accumulator) — no atomics, fixed summation order;
//   * warps are summed in warp order into one partial row per piece, the pieces of a slab in
//     piece order (an integer ticket picks the CTA that does that final sum; the order of the
//     sum itself is fixed), and y goes out through YDst (all ranks' buffers when sharded).
// AWSP addresses segments through a 32-bit per-row table, TCSR through 32-bit per-tile plus
// 16-bit in-tile offsets (the reference's blk_idx, tcsr.cpp:13,34, made two-level).
#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "common.cuh"
#include "plan.hpp"

namespace spmv {

namespace {

#ifndef SPMV_PANEL_WARPS
#define SPMV_PANEL_WARPS 8
#endif
constexpr int kPanelMaxWarps = SPMV_PANEL_WARPS;
constexpr int kPanelThreads = kPanelMaxWarps * 32;   // launch bound; the CTA size is a plan parameter
// chunks in flight per warp (cp.async groups)
#ifndef SPMV_PANEL_STAGES
#define SPMV_PANEL_STAGES 8
#endif
constexpr int kRingStages = SPMV_PANEL_STAGES;
// Multi-row mode is bound by the pass-by-pass retire latency, not by bytes in flight: a 4-deep ring
// leaves room for a second CTA per SM at 2048-column slabs (config-5 slab: 182.5 us with 8 warps per
// SM, 160.6 us with 16; profiles/r01_notes.md)
constexpr int kMrStages = 4;
#ifndef SPMV_LOB_STAGES
#define SPMV_LOB_STAGES 4
#endif
constexpr int kLobStages = SPMV_LOB_STAGES;    // lane-owned blocks: room for the three x buffers next to two CTAs per SM
template <bool MR, bool LOB = false> constexpr int stages_of() { return LOB ? kLobStages : MR ? kMrStages : kRingStages; }

template <int IDXB> struct ColIdx;
template <> struct ColIdx<8> {
    using Vec = uint32_t;                    // 4 x u8
    static __device__ __forceinline__ void copy(Vec *dst, const void *base, uint32_t g, bool ok)
    {
        cp_async4_zfill(dst, reinterpret_cast<const Vec *>(base) + g, ok);
    }
    static __device__ __forceinline__ void unpack(Vec v, uint32_t (&c)[4])
    {
        c[0] = v & 0xffu; c[1] = (v >> 8) & 0xffu; c[2] = (v >> 16) & 0xffu; c[3] = v >> 24;
    }
    static __device__ __forceinline__ Vec zero() { return 0u; }
};
template <> struct ColIdx<16> {
    using Vec = uint2;                       // 4 x u16
    static __device__ __forceinline__ void copy(Vec *dst, const void *base, uint32_t g, bool ok)
    {
        cp_async8_zfill(dst, reinterpret_cast<const Vec *>(base) + g, ok);
    }
    static __device__ __forceinline__ void unpack(Vec v, uint32_t (&c)[4])
    {
        c[0] = v.x & 0xffffu; c[1] = v.x >> 16; c[2] = v.y & 0xffffu; c[3] = v.y >> 16;
    }
    static __device__ __forceinline__ Vec zero() { return make_uint2(0u, 0u); }
};

// per-warp shared memory: [acc: W floats][vals ring: kStages x 32 float4][idx ring][list: 128 x uint4]
//                         [slot info][multi-row mode: per-lane (x, row slot) ring]
constexpr int kListCap = 128;              // rows a warp can take from one metadata batch
constexpr int kMetaBatch = 8;              // 32-row blocks of metadata fetched together
// B = 2 (batched form, one row per chunk only): two accumulator rows per warp and a second x per ring slot
template <int IDXB, int kStages, bool MR, bool LOB = false, int B = 1> __host__ __device__ constexpr int warp_smem_bytes(int W)
{
    return W * 4 * B + kStages * 32 * 16 + kStages * 32 * (IDXB == 8 ? 4 : 8) + (LOB ? 0 : kListCap * 16) + kStages * 8 +
           (MR ? kStages * 32 * 8 : 0) + (B > 1 ? kStages * 8 : 0);
}

// Balanced flat decomposition.  The (slab, row) pairs, slab-major, form one sequence of
// T = slabs*M units; CTA c of G owns units [c*T/G, (c+1)*T/G): equal work for every CTA whatever
// the slab count, one resident wave.  A range that crosses a slab boundary is processed as
// consecutive *pieces* (slab, row range); every piece ends with a fixed-order sum of the CTA's
// warps into one partial row, and the last piece to arrive for a slab (integer ticket) adds that
// slab's partial rows in CTA order.  The decomposition depends only on (shape, G), so a plan
// always reproduces its results bit for bit.
__device__ __forceinline__ long long range_begin(long long c, long long T, long long G) { return c * T / G; }
__device__ __forceinline__ long long cta_of_unit(long long u, long long T, long long G) { return ((u + 1) * G - 1) / T; }

// LOB: lane-owned blocks (formats.hpp) — the units of the flat sequence are (slab, block) pairs
// (`units` = blocks per slab), lane l owns the columns congruent to l modulo 32.
// B = 2: batched form — two activation vectors x[0], x[1] (row stride ldx) against the same
// A: a row segment is streamed once if EITHER vector is active there and used for both (a vector with
// x == 0 adds an exact zero), so every y[b] is bit-identical to a single-vector call.
template <int IDXB, bool TILED, bool MR, int kStages, bool LOB = false, int B = 1>
__global__ void __launch_bounds__(kPanelThreads)
panel_kernel(const float4 *__restrict__ vals, const void *__restrict__ idx,
             const uint32_t *__restrict__ off, const uint16_t *__restrict__ rel,
             const float *__restrict__ x, const YDst yd, float *__restrict__ partial,
             unsigned *__restrict__ tickets, int M, int N, int W, int row_blocks, int slabs, int kmax,
             int units, int block_rows, int cbits, long long ldx, long long ldy)
{
    static_assert(B == 1 || (B == 2 && !MR && !LOB), "the batched form exists for one-row-per-chunk plans");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int last_flag;
    using CI = ColIdx<IDXB>;
    using IVec = typename CI::Vec;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_warps = blockDim.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    unsigned char *wbase = smem_raw + (size_t)warp * warp_smem_bytes<IDXB, kStages, MR, LOB, B>(W);
    float *acc = reinterpret_cast<float *>(wbase);        // B rows of W accumulators
    float4 *ring_v = reinterpret_cast<float4 *>(wbase + (size_t)W * 4 * B);
    IVec *ring_i = reinterpret_cast<IVec *>(wbase + (size_t)W * 4 * B + kStages * 32 * 16);
    uint4 *list = reinterpret_cast<uint4 *>(wbase + (size_t)W * 4 * B + kStages * 32 * 16 + kStages * 32 * sizeof(IVec));
    uint2 *sinfo = reinterpret_cast<uint2 *>(list + (LOB ? 0 : kListCap));   // per ring slot: (valid lanes, x of the row | passes)
    uint2 *ring_m = sinfo + kStages;                             // MR: per lane (x of its row, row slot in the chunk)
    float *sinfo_x1 = reinterpret_cast<float *>(sinfo + kStages);             // B = 2: the second vector's x of the slot's row
    const int wg = blockIdx.x * n_warps + warp;           // trace id
    (void)wg;
    SPMV_STAMP(wg, 0);
    SPMV_STAMP_SMID(wg, 8);

    pdl_wait();
    const long long T = (long long)slabs * units, G = gridDim.x;
    const long long u_begin = range_begin(blockIdx.x, T, G), u_end = range_begin(blockIdx.x + 1, T, G);

    // ---- metadata of one 32-row block of a slab: lane = row ---------------------------------------
    struct Meta { float xv, xv1; uint32_t g0, g1; };
    // (Round 2 A/B: visiting a slab's 32-row blocks with a stride coprime to the block count, so that a
    // CTA's rows are spread over the whole slab, was SLOWER — config 2 awsp 21.3 vs 20.0 us, config 3
    // 10.8 vs 9.9 — so the 1.5x spread between SMs is not an address effect; contiguous ranges stay.)
    auto load_meta = [&](int slab, int rb) {
        Meta m;
        const int row = rb * 32 + lane;
        m.xv = row < M ? __ldg(x + row) : 0.0f;
        m.xv1 = (B > 1 && row < M) ? __ldg(x + ldx + row) : 0.0f;
        if (TILED) {
            const size_t t = (size_t)slab * (row_blocks + 1) + rb;
            const uint32_t tb = __ldg(off + t), te = __ldg(off + t + 1);
            const uint32_t r = __ldg(rel + ((size_t)slab * row_blocks + rb) * 32 + lane);
            const uint32_t rn = __shfl_down_sync(kFull, r, 1);
            m.g0 = tb + r;
            m.g1 = lane < 31 ? tb + rn : te;
        } else {
            const size_t o = (size_t)slab * ((size_t)M + 1) + row;
            m.g0 = row < M ? __ldg(off + o) : 0u;
            m.g1 = row < M ? __ldg(off + o + 1) : 0u;
        }
        return m;
    };

    // ---- kStages chunks in flight: cp.async ring, one commit group per chunk -------------------
    // A warp's rows form one flat sequence of chunks (32 groups) to the ring: `retire` finishes
    // the oldest slot (read-modify-write of the accumulator row), then the slot is refilled with
    // the chunk at hand.  Branch-free per chunk: lanes past the segment's end zero-fill their
    // slots, compute like everyone and only their stores are predicated off, so an idle lane can
    // never overwrite a live lane's update.  Slot info (valid lanes, x of the row) sits next to
    // the ring; a piece starts with everything zeroed, so warm-up and drain are the same code.
    int it = 0;
    auto retire = [&](int s) {
        const uint2 info = sinfo[s];
        const float4 a = ring_v[s * 32 + lane];
        uint32_t c[4];
        CI::unpack(ring_i[s * 32 + lane], c);
        const float p = __uint_as_float(info.y);
        // the four columns of a group are distinct (pads use an absent column)
        float r0 = acc[c[0]], r1 = acc[c[1]], r2 = acc[c[2]], r3 = acc[c[3]];
        r0 = fmaf(a.x, p, r0); r1 = fmaf(a.y, p, r1); r2 = fmaf(a.z, p, r2); r3 = fmaf(a.w, p, r3);
        if (lane < info.x) { acc[c[0]] = r0; acc[c[1]] = r1; acc[c[2]] = r2; acc[c[3]] = r3; }
        if (B > 1) {                                      // the same group against the second vector's accumulator row
            const float q = sinfo_x1[s];
            float *acc1 = acc + W;
            float t0 = acc1[c[0]], t1 = acc1[c[1]], t2 = acc1[c[2]], t3 = acc1[c[3]];
            t0 = fmaf(a.x, q, t0); t1 = fmaf(a.y, q, t1); t2 = fmaf(a.z, q, t2); t3 = fmaf(a.w, q, t3);
            if (lane < info.x) { acc1[c[0]] = t0; acc1[c[1]] = t1; acc1[c[2]] = t2; acc1[c[3]] = t3; }
        }
        __syncwarp();                                     // next chunk may be another row
    };
    auto run_list = [&](int n_rows) {
        for (int r = 0; r < n_rows; r++) {
            const uint4 m = list[r];
#pragma unroll 1
            for (uint32_t g = m.x; g < m.y; g += 32) {
                const int s = it++ & (kStages - 1);
                cp_async_wait<kStages - 1>();             // the oldest group (slot s) has landed
                retire(s);
                const uint32_t gg = g + lane;
                const bool ok = gg < m.y;
                const uint32_t gs = ok ? gg : g;
                cp_async16_zfill(ring_v + s * 32 + lane, vals + gs, ok);
                CI::copy(ring_i + s * 32 + lane, idx, gs, ok);
                cp_async_commit();
                if (lane == 0) {
                    sinfo[s] = make_uint2(min(32u, m.y - g), m.z);
                    if (B > 1) sinfo_x1[s] = __uint_as_float(m.w);
                }
            }
        }
    };

    // Multi-row mode (short segments, e.g. 1 % density): a chunk is 32 consecutive groups of the
    // warp's active rows laid end to end, so every cp.async moves 32 full lanes however short
    // the rows are.  Rows of one chunk may share columns, so the chunk is retired in passes, one
    // row (slot) at a time, in ascending row order.
    auto retire_mr = [&](int s) {
        const uint2 info = sinfo[s];                      // (valid lanes, passes)
        const float4 a = ring_v[s * 32 + lane];
        uint32_t c[4];
        CI::unpack(ring_i[s * 32 + lane], c);
        const uint2 mine = ring_m[s * 32 + lane];         // written by this lane at issue time
        const float p = __uint_as_float(mine.x);
        const bool live = lane < info.x;
        for (uint32_t pass = 0; pass < info.y; pass++) {
            if (live && mine.y == pass && p != 0.0f) {
                float r0 = acc[c[0]], r1 = acc[c[1]], r2 = acc[c[2]], r3 = acc[c[3]];
                r0 = fmaf(a.x, p, r0); r1 = fmaf(a.y, p, r1); r2 = fmaf(a.z, p, r2); r3 = fmaf(a.w, p, r3);
                acc[c[0]] = r0; acc[c[1]] = r1; acc[c[2]] = r2; acc[c[3]] = r3;
            }
            __syncwarp();
        }
    };
    auto run_list_mr = [&](int n_rows) {
        for (int base = 0; base < n_rows; base += 32) {   // lane i <-> row base + i of the list
            const uint4 m = base + lane < n_rows ? list[base + lane] : make_uint4(0u, 0u, 0u, 0u);
            const int cgr = (int)(m.y - m.x);             // groups of this lane's row (>= 1 when present)
            const int pend = warp_incl_scan(cgr, lane);   // flat position one past the row's last group
            const int pbeg = pend - cgr;
            const int total = __shfl_sync(kFull, pend, 31);
#pragma unroll 1
            for (int k0 = 0; k0 < total; k0 += 32) {
                const int s = it++ & (kStages - 1);
                cp_async_wait<kStages - 1>();
                retire_mr(s);
                // which row does flat position k0 + lane belong to?  rows ending inside the chunk
                // set one bit each; rows that ended before it are counted by a ballot
                const unsigned e = (unsigned)(pend - k0 - 1);
                const unsigned endbit = (cgr > 0 && e < 32u) ? (1u << e) : 0u;
                const unsigned ends = __reduce_or_sync(kFull, endbit);
                // passes are numbered over the rows that take part in the arithmetic (x != 0)
                const unsigned ends_act = __reduce_or_sync(kFull, __uint_as_float(m.z) != 0.0f ? endbit : 0u);
                const int i_first = __popc(__ballot_sync(kFull, cgr > 0 && pend <= k0));
                const int q = k0 + lane;
                const bool ok = q < total;
                const int i = min(31, i_first + __popc(ends & lt));
                const uint32_t g0_i = __shfl_sync(kFull, m.x, i);
                const int pb_i = __shfl_sync(kFull, pbeg, i);
                const uint32_t x_i = __shfl_sync(kFull, m.z, i);
                const uint32_t gs = ok ? g0_i + (uint32_t)(q - pb_i) : g0_i;
                cp_async16_zfill(ring_v + s * 32 + lane, vals + gs, ok);
                CI::copy(ring_i + s * 32 + lane, idx, gs, ok);
                cp_async_commit();
                ring_m[s * 32 + lane] = make_uint2(x_i, (uint32_t)__popc(ends_act & lt));
                const int n_lanes = min(32, total - k0);
                if (lane == 0) sinfo[s] = make_uint2((uint32_t)n_lanes, (uint32_t)__popc(ends_act) + 1u);
            }
        }
    };
    // Lane-owned blocks: a chunk is 32 groups of one (slab, block); lane l's four entries belong to
    // columns 32*c + l, so lanes never meet in an accumulator and the chunk retires in one pass with
    // no votes, no predicates and no bank conflicts.  x enters as a multiplier: the CTA keeps the x
    // slices of three consecutive blocks in shared memory (the block being streamed, the one before
    // it, whose last chunks may still sit in a ring, and the next one, fetched into registers while
    // the current block streams).  All warps of the CTA walk the same blocks, each taking every
    // n_warps-th chunk.
    float *xbuf = reinterpret_cast<float *>(smem_raw + (size_t)n_warps * warp_smem_bytes<IDXB, kStages, MR, LOB, B>(W));
    auto retire_lob = [&](int s, uint32_t xo) {           // xo: where the chunk's block keeps its x slice
        const float4 a = ring_v[s * 32 + lane];
        uint32_t c[4];
        CI::unpack(ring_i[s * 32 + lane], c);
        const float *xb = xbuf + xo;
        const uint32_t cmask = (1u << cbits) - 1u;
        const float p0 = xb[c[0] >> cbits], p1 = xb[c[1] >> cbits], p2 = xb[c[2] >> cbits], p3 = xb[c[3] >> cbits];
        float *al = acc + lane;
        // one after the other: two entries of a group may share a column (different rows)
        { float *q = al + ((c[0] & cmask) << 5); *q = fmaf(a.x, p0, *q); }
        { float *q = al + ((c[1] & cmask) << 5); *q = fmaf(a.y, p1, *q); }
        { float *q = al + ((c[2] & cmask) << 5); *q = fmaf(a.z, p2, *q); }
        { float *q = al + ((c[3] & cmask) << 5); *q = fmaf(a.w, p3, *q); }
    };
    constexpr int kXPerThread = kLobMaxBlockRows / kPanelThreads;   // block_rows <= 1024, 256 threads
    // The ring slots are visited in a fixed rotation (the loop body is unrolled over them), so what a
    // slot holds — nothing, or a chunk and the x slice that goes with it — lives in registers: no
    // slot table in shared memory, no warp-level synchronisation at all (a lane reads back only what
    // it copied itself and only touches its own accumulators).
    auto run_blocks = [&](int slab, int blk_a, int blk_b) {
        const uint32_t *ob = off + (size_t)slab * (units + 1);
        uint32_t g1 = __ldg(ob + blk_a), g2 = __ldg(ob + blk_a + 1);
        float nx[kXPerThread];
        auto fetch_x = [&](int blk) {                     // this thread's share of a block's x slice -> registers
#pragma unroll
            for (int k = 0; k < kXPerThread; k++) {
                const int i = tid + k * (int)blockDim.x;
                const long long row = (long long)blk * block_rows + i;
                nx[k] = (i < block_rows && row < M) ? __ldg(x + row) : 0.0f;
            }
        };
        auto store_x = [&](int buf) {
#pragma unroll
            for (int k = 0; k < kXPerThread; k++) {
                const int i = tid + k * (int)blockDim.x;
                if (i < block_rows) xbuf[buf * block_rows + i] = nx[k];
            }
        };
        bool live[kStages]; uint32_t xo[kStages];
#pragma unroll
        for (int s = 0; s < kStages; s++) { live[s] = false; xo[s] = 0u; }
        __syncthreads();                                  // every warp has left the previous piece
        fetch_x(blk_a);
        store_x(0);
        __syncthreads();
        const uint32_t stride = 32u * n_warps;
        for (int blk = blk_a; blk < blk_b; blk++) {
            const int buf = (blk - blk_a) % 3;
            const uint32_t xoff = (uint32_t)(buf * block_rows);
            const uint32_t g0 = g1;
            g1 = g2;
            if (blk + 2 <= blk_b) g2 = __ldg(ob + min(blk + 2, units));   // one block ahead
            if (blk + 1 < blk_b) fetch_x(blk + 1);
            // One turn of the rotation retires everything issued before it, so at the end of a
            // block only this block's chunks are in flight — a turn is made even when the warp
            // has no chunk here — and only the current and the previous block's x slices are
            // ever live when the third buffer is overwritten.
            uint32_t g = g0 + 32u * warp;
            bool first = true;
            while (first || g < g1) {
                first = false;
#pragma unroll
                for (int s = 0; s < kStages; s++) {
                    cp_async_wait<kStages - 1>();
                    if (live[s]) retire_lob(s, xo[s]);
                    const bool have = g < g1;
                    if (have) {
                        cp_async16(ring_v + s * 32 + lane, vals + g + lane);
                        CI::copy(ring_i + s * 32 + lane, idx, g + lane, true);
                        g += stride;
                    }
                    cp_async_commit();
                    live[s] = have; xo[s] = xoff;
                }
            }
            if (blk + 1 < blk_b) {
                store_x((buf + 1) % 3);
                __syncthreads();
            }
        }
        cp_async_wait<0>();
#pragma unroll
        for (int s = 0; s < kStages; s++)
            if (live[s]) retire_lob(s, xo[s]);
    };
    int piece = 0;
    for (long long u = u_begin; u < u_end; piece++) {
        const int slab = (int)(u / units);
        const int row_a = (int)(u - (long long)slab * units);
        const int row_b = (int)min((long long)units, row_a + (u_end - u));
        u += row_b - row_a;

        for (int c = lane; c < W * B; c += 32) acc[c] = 0.0f;
        for (int k = lane; k < kStages * 32; k += 32) ring_i[k] = CI::zero();
        if (lane < kStages) sinfo[lane] = make_uint2(0u, 0u);
        __syncwarp();

        // activation compaction: every warp looks at every 32-row block of the piece (lane =
        // row): ballot over "in range, x != 0, segment non-empty", order-preserving popc rank;
        // warp w keeps the active rows whose rank is w modulo the warp count, so the warps of
        // a CTA end up with equal shares whatever the piece length.
        // Metadata is fetched kMetaBatch blocks at a time, one batch ahead of its use, so its
        // latency is paid once per 128 rows and overlaps the previous batch's streaming.
        // Pieces with at least one block per warp: warp w owns blocks w, w + n_warps, ... outright
        // (no redundant metadata loads, one metadata latency per 256 of its rows, statistically
        // balanced).  Shorter pieces: every warp scans every block (a single batch) and keeps the
        // active rows whose rank is w modulo the warp count (exact balance).
        if (LOB) {
            run_blocks(slab, row_a, row_b);
        } else {
        const int blk_a = row_a >> 5, blk_b = (row_b + 31) >> 5;
        const bool by_block = (blk_b - blk_a) >= n_warps;
        const int bstep = by_block ? n_warps : 1;
        const int bfirst = blk_a + (by_block ? warp : 0);
        int rank_base = 0;
        for (int blk0 = bfirst; blk0 < blk_b; blk0 += kMetaBatch * bstep) {
            // metadata of kMetaBatch blocks at once (one latency per 256 rows; a short piece
            // is covered by a single batch), consumed as two lists of four blocks
            Meta cur[kMetaBatch];
#pragma unroll
            for (int j = 0; j < kMetaBatch; j++)
                if (blk0 + j * bstep < blk_b) cur[j] = load_meta(slab, blk0 + j * bstep);
#pragma unroll
            for (int half = 0; half < 2; half++) {
                if (blk0 + half * (kMetaBatch / 2) * bstep >= blk_b) break;
                int cnt = 0;
#pragma unroll
                for (int jj = 0; jj < kMetaBatch / 2; jj++) {
                    const int j = half * (kMetaBatch / 2) + jj;
                    const int blk = blk0 + j * bstep;
                    const int row = blk * 32 + lane;
                    const bool valid = blk < blk_b && row >= row_a && row < row_b && (cur[j].xv != 0.0f || (B > 1 && cur[j].xv1 != 0.0f)) &&
                                       cur[j].g1 > cur[j].g0;
                    const unsigned mask = __ballot_sync(kFull, valid);
                    const int rank = rank_base + __popc(mask & lt);
                    const bool mine = valid && (by_block || (rank & (n_warps - 1)) == warp);   // n_warps is a power of two
                    const unsigned mm = __ballot_sync(kFull, mine);
                    if (mine) list[cnt + __popc(mm & lt)] = make_uint4(cur[j].g0, cur[j].g1, __float_as_uint(cur[j].xv), __float_as_uint(cur[j].xv1));
                    cnt += __popc(mm);
                    rank_base += __popc(mask);
                }
                __syncwarp();
                if (piece == 0 && blk0 == bfirst && half == 0) SPMV_STAMP(wg, 1);
                if (MR) run_list_mr(cnt); else run_list(cnt);
                __syncwarp();                             // the list is rewritten next
            }
        }
        }
        SPMV_STAMP(wg, 2);
        cp_async_wait<0>();
        SPMV_STAMP(wg, 3);
#pragma unroll 1
        for (int k = 0; k < kStages && !LOB; k++) {       // (run_blocks drains its own ring)
            if (MR) retire_mr(it++ & (kStages - 1)); else retire(it++ & (kStages - 1));
        }
        SPMV_STAMP(wg, 4);

        // ---- fixed-order sum over warps, then over the slab's pieces -----------------------------
        __syncthreads();
        SPMV_STAMP(wg, 5);
        const long long s_begin = (long long)slab * units;
        const long long c_lo = cta_of_unit(s_begin, T, G), c_hi = cta_of_unit(s_begin + units - 1, T, G);
        const int n_pieces = (int)(c_hi - c_lo + 1);
        const int col0 = slab * W;
        const int n_valid = min(W, N - col0);
        const int wstride = warp_smem_bytes<IDXB, kStages, MR, LOB, B>(W) / 4;
        const float *acc0 = reinterpret_cast<const float *>(smem_raw);
        // vector b's partial rows live at [b][CTA * kmax + piece]; its y at row b of the batch (stride ldy)
        const size_t rows_per_b = (size_t)gridDim.x * kmax;
#pragma unroll
        for (int b = 0; b < B; b++) {
            float *dst = partial + ((size_t)b * rows_per_b + (size_t)blockIdx.x * kmax + piece) * W;
            for (int c = tid; c < n_valid; c += blockDim.x) {
                float s = acc0[b * W + c];
                for (int w = 1; w < n_warps; w++) s += acc0[(size_t)w * wstride + b * W + c];
                if (n_pieces == 1) y_store(yd, (size_t)b * ldy + col0 + c, s); else dst[c] = s;
            }
        }
        SPMV_STAMP(wg, 6);
        if (n_pieces > 1) {
            // partial row j of this slab was written by CTA c_lo + j as its piece number
            // (slab - first slab of that CTA); the row offsets go to shared memory once
            float4 *scratch = reinterpret_cast<float4 *>(smem_raw);          // accumulators are dead by now
            uint32_t *row_of = reinterpret_cast<uint32_t *>(scratch + blockDim.x);
            __syncthreads();
            for (int j = tid; j < n_pieces; j += blockDim.x) {
                const long long c = c_lo + j;
                row_of[j] = (uint32_t)(c * kmax + (slab - (int)(range_begin(c, T, G) / units)));
            }
            // (split_reduce_rows starts with a barrier, which also publishes row_of)
            for (int b = 0; b < B; b++) {
                const float *pb = partial + (size_t)b * rows_per_b * W;
                split_reduce_rows(yd, (size_t)b * ldy + col0, [&](int j) { return pb + (size_t)row_of[j] * W; },
                                  &tickets[(size_t)b * slabs + slab], n_pieces, W, n_valid, &last_flag, scratch);
                if (B > 1) __syncthreads();               // the scratch is reused by the next vector
            }
        }
        __syncthreads();                                  // shared memory is reused by the next piece
        SPMV_STAMP(wg, 7);
    }
}

template <int IDXB, bool TILED, bool MR, bool LOB = false>
int launch_variant(spmv_plan *p, const float *x, const YDst &y, cudaStream_t st)
{
    auto k = panel_kernel<IDXB, TILED, MR, stages_of<MR, LOB>(), LOB>;
    static int smem_set[16] = {0};                        // per device: largest dynamic smem opted in so far
    if (p->smem > 48 * 1024 && p->device >= 0 && p->device < 16 && smem_set[p->device] < p->smem) {
        SPMV_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, p->smem));
        smem_set[p->device] = p->smem;
    }
    const DevPanel &d = p->panel;
    int cbits = 0;
    while ((32 << cbits) < d.slab_cols) cbits++;
    SPMV_CUDA(launch_k(k, p->grid, dim3(p->block), p->smem, st, reinterpret_cast<const float4 *>(d.vals), d.idx, d.off, d.rel,
                       x, y, p->partial, p->tickets, (int)p->M, (int)p->N, d.slab_cols, d.row_blocks, d.slabs, d.kmax,
                       LOB ? d.lob_blocks : (int)p->M, d.block_rows, cbits, 0LL, 0LL));
    return SPMV_OK;
}

// Two vectors per pass (one-row-per-chunk plans): a 4-deep ring so that two accumulator rows per warp
// still leave room for two CTAs per SM; same grid, same decomposition, so y[b] is bit-identical to
// a single-vector call.
template <int IDXB, bool TILED>
int launch_batch2(spmv_plan *p, const float *x, long long ldx, const YDst &y, long long ldy, cudaStream_t st)
{
    auto k = panel_kernel<IDXB, TILED, false, kMrStages, false, 2>;
    const DevPanel &d = p->panel;
    const int smem = d.warps * warp_smem_bytes<IDXB, kMrStages, false, false, 2>(d.slab_cols);
    if (smem > (p->max_smem_optin > 0 ? p->max_smem_optin : 227 * 1024) || 2 * smem + 2048 > 228 * 1024) return SPMV_ERR_UNSUPPORTED;
    static int smem_set[16] = {0};
    if (smem > 48 * 1024 && p->device >= 0 && p->device < 16 && smem_set[p->device] < smem) {
        SPMV_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        smem_set[p->device] = smem;
    }
    SPMV_CUDA(launch_k(k, p->grid, dim3(p->block), (size_t)smem, st, reinterpret_cast<const float4 *>(d.vals), d.idx, d.off, d.rel,
                       x, y, p->partial, p->tickets, (int)p->M, (int)p->N, d.slab_cols, d.row_blocks, d.slabs, d.kmax,
                       (int)p->M, d.block_rows, 0, ldx, ldy));
    return SPMV_OK;
}

} // namespace

int launch_panel(spmv_plan *p, const float *d_x, const YDst &d_y, cudaStream_t st)
{
    if (p->N == 0) return SPMV_OK;
    if (p->M == 0) {                                      // no rows: y = 0 (plain stores; the caller joins)
        for (int k = 0; k < d_y.n; k++) SPMV_CUDA(cudaMemsetAsync(d_y.p[k], 0, (size_t)p->N * sizeof(float), st));
        return SPMV_OK;
    }
    const DevPanel &d = p->panel;
    if (d.rs_grid > 0) return launch_panel_rs(p, d_x, d_y, st);           // one row per chunk: the register-staged form
    if (d.block_rows > 0) return launch_variant<16, false, false, true>(p, d_x, d_y, st);
    if (d.multirow) {
        if (d.index_bits == 8) return d.tiled ? launch_variant<8, true, true>(p, d_x, d_y, st) : launch_variant<8, false, true>(p, d_x, d_y, st);
        return d.tiled ? launch_variant<16, true, true>(p, d_x, d_y, st) : launch_variant<16, false, true>(p, d_x, d_y, st);
    }
    if (d.index_bits == 8) return d.tiled ? launch_variant<8, true, false>(p, d_x, d_y, st) : launch_variant<8, false, false>(p, d_x, d_y, st);
    return d.tiled ? launch_variant<16, true, false>(p, d_x, d_y, st) : launch_variant<16, false, false>(p, d_x, d_y, st);
}

// Batched form: B = 2 vectors in one pass over A's row segments (awsp / tcsr, one row per chunk);
// returns SPMV_ERR_UNSUPPORTED otherwise (the caller then runs the vectors one by one).
int launch_panel_batch(spmv_plan *p, const float *d_x, long long ldx, const YDst &d_y, long long ldy, int B, cudaStream_t st)
{
    const DevPanel &d = p->panel;
    if (B != 2 || d.multirow || d.block_rows > 0 || p->M == 0 || p->N == 0 || !p->panel.batch_ok) return SPMV_ERR_UNSUPPORTED;
    if (d.index_bits == 8) return d.tiled ? launch_batch2<8, true>(p, d_x, ldx, d_y, ldy, st) : launch_batch2<8, false>(p, d_x, ldx, d_y, ldy, st);
    return d.tiled ? launch_batch2<16, true>(p, d_x, ldx, d_y, ldy, st) : launch_batch2<16, false>(p, d_x, ldx, d_y, ldy, st);
}

// Geometry: a 1-D grid of G CTAs over the flat (slab, row) sequence, one resident wave.
// Every CTA pays a few microseconds of serial latency (metadata, first chunks, cross-warp and
// cross-piece sums), so more than one wave only adds latency (measured, profiles/r01_notes.md);
// 8-warp CTAs, two per SM where shared memory allows (a third adds more fixed cost than it hides:
// 444 CTAs 21.2 us vs 296 CTAs 16.4 us on the 4096x4096 / 50 % case).
int configure_panel(spmv_plan *p, const HostPanel &h, const spmv_options_t *o)
{
    DevPanel &d = p->panel;
    d.slab_cols = h.slab_cols; d.index_bits = h.index_bits; d.slabs = h.slabs;
    d.row_blocks = h.row_blocks; d.tiled = h.tiled;
    d.block_rows = h.block_rows; d.lob_blocks = h.lob_blocks;
    const bool lob = h.block_rows > 0;
    // short segments (fewer than 6 groups on average, i.e. chunks less than a fifth full): pack
    // several rows into one chunk
    const double segs = std::max<double>(1.0, (double)h.nonempty_segments);
    // (round 2: with the register-staged one-row-per-chunk kernel the crossover moved down — config 1, 6.4 groups
    // per segment: 9.25 us there against 10.26 us multi-row)
    d.multirow = (double)h.groups / segs < 6.0;
    if (o && o->chunk_mode == 1) d.multirow = false;
    if (o && o->chunk_mode == 2) d.multirow = true;
    if (lob) d.multirow = false;
    const int smem_cap = p->max_smem_optin > 0 ? p->max_smem_optin : 227 * 1024;
    const int per_warp = lob ? warp_smem_bytes<16, kLobStages, false, true>(h.slab_cols) : d.multirow ? (h.index_bits == 8 ? warp_smem_bytes<8, kMrStages, true>(h.slab_cols) : warp_smem_bytes<16, kMrStages, true>(h.slab_cols))
                                    : (h.index_bits == 8 ? warp_smem_bytes<8, kRingStages, false>(h.slab_cols) : warp_smem_bytes<16, kRingStages, false>(h.slab_cols));
    int warps = kPanelMaxWarps;
    if (o && o->warps_per_col > 0) {
        warps = 1;
        while (warps * 2 <= std::min(kPanelMaxWarps, o->warps_per_col)) warps *= 2;   // power of two
    }
    const int xbuf_bytes = lob ? 3 * h.block_rows * 4 : 0;               // lane-owned blocks: three x slices per CTA
    if (lob) warps = kPanelMaxWarps;                                      // the x slices are staged by 256 threads
    while (!lob && warps > 1 && warps * per_warp > smem_cap) warps /= 2;
    if (warps * per_warp + xbuf_bytes > smem_cap)
        return set_error(SPMV_ERR_UNSUPPORTED, "panel: %d bytes of shared memory exceed the device limit", warps * per_warp + xbuf_bytes);
    d.warps = warps;
    p->block = warps * 32;
    p->smem = warps * per_warp + xbuf_bytes;
    p->tile_width = h.slab_cols;
    p->col_tiles = h.slabs;
    p->kernels_per_run = 1;

    const int slabs = std::max(1, h.slabs);
    const int64_t M = std::max<int64_t>(1, lob ? h.lob_blocks : h.M);      // units per slab of the flat sequence
    const int resident = std::max(1, std::min(std::min(lob ? 3 : 2, 2048 / p->block), (228 * 1024) / (p->smem + 1024)));
    int64_t G = (int64_t)p->sm_count * resident;
    // a whole number of CTAs per slab when that costs under 10 % of the grid: no piece crosses a
    // slab boundary, so no CTA pays the per-piece overhead twice (288 vs 296 CTAs: 11.9 vs 13.8 us)
    if (G >= slabs && (G / slabs) * slabs * 10 >= G * 9) G = (G / slabs) * slabs;
    if (o && o->row_splits > 0) G = (int64_t)slabs * o->row_splits;       // forced: row_splits CTAs per slab
    const int64_t T = (int64_t)slabs * M;
    G = std::max<int64_t>(1, std::min<int64_t>(G, lob ? T : (T + 31) / 32));   // at least 32 rows (one block) per CTA
    // the final sum stages one row offset per piece of a slab behind blockDim float4 of scratch: keep
    // the pieces per slab inside the CTA's shared memory (forced options can ask for thousands)
    const int64_t max_pieces = std::max<int64_t>(1, ((int64_t)p->smem - (int64_t)p->block * 16) / 4 - 2);
    if ((G + slabs - 1) / slabs + 1 > max_pieces) G = std::max<int64_t>(1, (max_pieces - 1) * slabs);
    const int64_t max_range = (T + G - 1) / G;
    d.kmax = (int)(2 + max_range / M);
    p->row_splits = (int)((G + slabs - 1) / slabs);                        // reported: CTAs per slab (rounded up)
    p->grid = dim3((unsigned)G, 1, 1);
    // scratch: one partial row per (CTA, piece) + one ticket per slab
    p->partial = nullptr; p->tickets = nullptr;
    d.batch_ok = !lob && !d.multirow;                     // room for the second vector of the batched form
    const size_t copies = d.batch_ok ? 2 : 1;
    int rc = alloc_panel_scratch(p, copies * (size_t)G * d.kmax * h.slab_cols, copies * (size_t)slabs);
    // Single-vector calls on one-row-per-chunk plans take the register-staged form (panel_rs.cu) unless the
    // caller forced this kernel's geometry; the ring kernel above stays for the pair form, the forced
    // options, the multi-row and the lane-owned modes.  SPMV_PANEL_RS=0: development switch (same-box A/B).
    d.rs_grid = 0;
    const char *e = std::getenv("SPMV_PANEL_RS");
    const bool forced = o && (o->row_splits > 0 || o->warps_per_col > 0);
    if (!rc && !lob && !d.multirow && !forced && !(e && std::atoi(e) == 0) && h.M > 0 && h.N > 0) {
        rc = configure_panel_rs(p, h);
        if (!rc && d.rs_grid > 0) p->kernels_per_run = 2;
    }
    return rc;
}

} // namespace spmv
