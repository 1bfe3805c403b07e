// panel_rs.cu — the row-panel format (formats.hpp: HostPanel) for the dense-ish shapes, one row per
// 32-group chunk, with the chunks in flight held in REGISTERS.
//
// Same variants and the same HBM format as panel.cu (awsp: awsp_kernel_v0/v1/v2, awsp.cu:5-317,
// awsp_ref_kernel, awsp_ref.cu:6-185; tcsr: csr_tiling_kernel, csr_tiling.cu:24-114); what changes is
// how a CTA spends its time.  Round 1's kernel (panel.cu) moves every chunk through a cp.async ring
// — two LSU passes per byte (copy-in + read-back) — runs 16 warps per SM, and every piece pays a
// head (block-wide list building) and a tail (partial row, fence, ticket, last-arriver sum) of ~2 us
// each on a 20 us kernel.  Round 2's measurements (profiles/r02_notes.md) say: a ring costs LSU
// wavefronts twice, loads in flight decide the rate, a launch costs 0.55 us.  Hence:
//   * kRsDepth = 8 chunks in flight per warp as plain 128-bit + 64-bit loads into registers (volatile asm
//     pins the issue point), 8-warp CTAs, two per SM — measured against 3 and 4 CTAs per SM with 4-6
//     chunks in flight (profiles/r02_notes.md): the deeper pipeline of fewer warps wins, and fewer
//     CTAs mean fewer partial rows;
//   * no block-wide barrier before the end of a piece: every warp scans the piece's 32-row blocks
//     itself (lane = row: x, segment range; ballot over "x != 0 and segment non-empty"; the x != 0.0f
//     test of awsp.cu:98,127) and keeps the active rows whose rank is its own modulo 8;
//   * the chunk pipeline runs on across list refills; products go to the warp's private accumulator
//     row in shared memory (distinct columns inside a segment, ascending rows per warp);
//   * a piece ends with one barrier, the fixed-order sum of the 8 warps, and ONE partial row in global
//     memory — no fence, no ticket; panel_rs_reduce_kernel (dependent launch, all SMs) adds a slab's
//     partial rows in CTA order and stores y through YDst.
// Deterministic: no atomics at all; the decomposition depends only on (shape, grid).
#include <algorithm>

#include "common.cuh"
#include "plan.hpp"

namespace spmv {

namespace {

constexpr int kRsWarps = 8;
constexpr int kRsThreads = kRsWarps * 32;
#ifndef SPMV_RS_DEPTH
#define SPMV_RS_DEPTH 8
#endif
constexpr int kRsDepth = SPMV_RS_DEPTH;       // chunks (32 groups = 768 B) in flight per warp
#ifndef SPMV_RS_BATCH
#define SPMV_RS_BATCH 8
#endif
constexpr int kRsBatch = SPMV_RS_BATCH;       // 32-row blocks scanned per list refill (their metadata lives in registers)
#ifndef SPMV_RS_CTAS
#define SPMV_RS_CTAS 2
#endif
// Bit 0: the reduce kernel releases its dependents at entry, and a CTA of the main kernel fetches the static
// metadata (segment offsets) of its first blocks BEFORE the grid dependency wait — so across back-to-back
// calls the next call's CTAs are resident, zeroed and one memory round trip ahead while the previous call's
// partial rows are still being summed.  Bit 1: the main kernel also releases the reduce kernel after its
// streaming loop.  Same-box A/B, us per call, awsp c2 / c0 / c1 / c3 then tcsr: off 16.56 11.62 8.92 9.78 |
// 17.62 12.81 10.10 10.23; bit 0: 16.06 11.31 8.56 9.41 | 16.98 11.69 9.32 9.32; bit 1: 17.06 11.98 9.08
// 10.06 | 18.06 13.15 10.22 10.47; both: 16.64 11.81 9.02 9.98 | 16.99 11.87 9.28 9.71.  Bit 0 it is.
#ifndef SPMV_RS_EARLY
#define SPMV_RS_EARLY 1
#endif
constexpr bool kRsEarly = (SPMV_RS_EARLY & 1) != 0;        // reduce kernel releases at entry, main kernel prefetches before its wait
constexpr bool kRsTrigger = (SPMV_RS_EARLY & 2) != 0;      // main kernel releases the reduce kernel after its streaming loop
constexpr int kRsCtas = SPMV_RS_CTAS;         // CTAs per SM the register budget is cut for (2: up to 128 registers per thread)
// rows a warp can get from one list refill: whole blocks when it owns them (BB), else its rank-interleaved share
__host__ __device__ constexpr int rs_list_rows(bool bb) { return bb ? kRsBatch * 32 : kRsBatch * 32 / kRsWarps; }

template <int IDXB> struct RsIdx;
template <> struct RsIdx<8> {
    using Vec = uint32_t;                     // 4 x u8
    static __device__ __forceinline__ Vec load(const void *base, uint32_t g, bool ok)
    {
        Vec r = 0u;
        if (ok) asm volatile("ld.global.nc.L1::no_allocate.b32 %0, [%1];" : "=r"(r) : "l"(reinterpret_cast<const Vec *>(base) + g));
        return r;
    }
    static __device__ __forceinline__ void unpack(Vec v, uint32_t (&c)[4])
    {
        c[0] = v & 0xffu; c[1] = (v >> 8) & 0xffu; c[2] = (v >> 16) & 0xffu; c[3] = v >> 24;
    }
};
template <> struct RsIdx<16> {
    using Vec = uint2;                        // 4 x u16
    static __device__ __forceinline__ Vec load(const void *base, uint32_t g, bool ok)
    {
        Vec r = make_uint2(0u, 0u);
        if (ok) asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(reinterpret_cast<const Vec *>(base) + g));
        return r;
    }
    static __device__ __forceinline__ void unpack(Vec v, uint32_t (&c)[4])
    {
        c[0] = v.x & 0xffffu; c[1] = v.x >> 16; c[2] = v.y & 0xffffu; c[3] = v.y >> 16;
    }
};

__device__ __forceinline__ float4 rs_load_vals(const float4 *p, bool ok)
{
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ok) asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}

__host__ __device__ constexpr int rs_warp_bytes(int W, bool bb) { return W * 4 + rs_list_rows(bb) * 16; }

// the flat (slab, row) sequence of T = slabs * M units cut into G equal ranges (as in panel.cu)
// (`small`: T * G fits 31 bits — every shape but the very largest — so the divisions are 32-bit ones; a 64-bit
// division is ~100 instructions and the reduce kernel has three of them on its critical path)
__device__ __forceinline__ long long rs_begin(long long c, long long T, long long G, bool small)
{
    return small ? (long long)((unsigned)c * (unsigned)T / (unsigned)G) : c * T / G;
}
__device__ __forceinline__ long long rs_owner(long long u, long long T, long long G, bool small)
{
    return small ? (long long)((((unsigned)u + 1u) * (unsigned)G - 1u) / (unsigned)T) : ((u + 1) * G - 1) / T;
}

// BB: the plan's pieces hold at least one 32-row block per warp, so a warp owns whole blocks (compiled as a
// separate instance: with the larger list and the ownership test in the code the short-piece case ran
// 15 % slower on config 2 — same-box A/B in profiles/r02_notes.md).
template <int IDXB, bool TILED, bool BB>
__global__ void __launch_bounds__(kRsThreads, kRsCtas)
panel_rs_kernel(const float4 *__restrict__ vals, const void *__restrict__ idx, const uint32_t *__restrict__ off,
                const uint16_t *__restrict__ rel, const float *__restrict__ x, float *__restrict__ partial,
                int M, int N, int W, int row_blocks, int slabs, int kmax, int small)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using IX = RsIdx<IDXB>;
    using IVec = typename IX::Vec;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned lt = (1u << lane) - 1u;
    float *acc = reinterpret_cast<float *>(smem_raw + (size_t)warp * rs_warp_bytes(W, BB));
    uint4 *list = reinterpret_cast<uint4 *>(acc + W);     // (first group, end group, x bits, -) of this warp's rows

    const int wg = blockIdx.x * kRsWarps + warp;          // (timeline builds only)
    SPMV_STAMP(wg, 0);
    SPMV_STAMP_SMID(wg, 8);
    if (!kRsEarly) pdl_wait();
    if (!kRsEarly) SPMV_STAMP(wg, 1);
    const long long T = (long long)slabs * M, G = gridDim.x;
    const long long u_begin = rs_begin(blockIdx.x, T, G, small), u_end = rs_begin(blockIdx.x + 1, T, G, small);

    // the chunks in flight: values, column ids, the row's x, valid lanes (0: empty slot)
    float4 sv[kRsDepth]; IVec si[kRsDepth]; float sx[kRsDepth]; int sn[kRsDepth];
#pragma unroll
    for (int k = 0; k < kRsDepth; k++) { sv[k] = make_float4(0.f, 0.f, 0.f, 0.f); si[k] = IVec(); sx[k] = 0.0f; sn[k] = 0; }

    auto consume = [&](const float4 &a, const IVec &iv, float p, int n) {
        if (n == 0) return;                               // (warp-uniform)
        uint32_t c[4];
        IX::unpack(iv, c);
        // the four columns of a group are distinct (pads use a column absent from the segment); idle lanes
        // hold zeros and only their stores are predicated off, so they cannot overwrite a live lane's update
        float r0 = acc[c[0]], r1 = acc[c[1]], r2 = acc[c[2]], r3 = acc[c[3]];
        r0 = fmaf(a.x, p, r0); r1 = fmaf(a.y, p, r1); r2 = fmaf(a.z, p, r2); r3 = fmaf(a.w, p, r3);
        if (lane < n) { acc[c[0]] = r0; acc[c[1]] = r1; acc[c[2]] = r2; acc[c[3]] = r3; }
        __syncwarp();                                     // the next chunk may be another row
    };

    int piece = 0;
    for (long long u = u_begin; u < u_end; piece++) {
        const int slab = small ? (int)((unsigned)u / (unsigned)M) : (int)(u / M);
        const int row_a = (int)(u - (long long)slab * M);
        const int row_b = (int)min((long long)M, row_a + (u_end - u));
        u += row_b - row_a;

        for (int c = lane; c < W; c += 32) acc[c] = 0.0f;
        __syncwarp();

        const int blk_a = row_a >> 5, blk_b = (row_b + 31) >> 5;
        int rank_base = 0;
        // metadata of up to kRsBatch blocks at once (one memory latency), lane = row; the NEXT batch's loads are
        // issued as soon as this batch's registers have been turned into the list, so their latency overlaps the
        // streaming of this batch's chunks
        float mx[kRsBatch]; uint32_t mg0[kRsBatch], mg1[kRsBatch];
        auto load_meta = [&](int blk0, bool statics = true, bool xs = true) {
            const int bstep_ = BB && (blk_b - blk_a) >= kRsWarps ? kRsWarps : 1;
#pragma unroll
            for (int j = 0; j < kRsBatch; j++) {
                const int rb = blk0 + j * bstep_, row = rb * 32 + lane;
                if (xs) mx[j] = 0.0f;
                if (statics) { mg0[j] = 0u; mg1[j] = 0u; }
                if (rb < blk_b && row < M) {
                    if (xs) mx[j] = __ldg(x + row);
                    if (!statics) continue;
                    if (TILED) {
                        const size_t t = (size_t)slab * (row_blocks + 1) + rb;
                        const uint32_t tb = __ldg(off + t), te = __ldg(off + t + 1);
                        const uint32_t r = __ldg(rel + ((size_t)slab * row_blocks + rb) * 32 + lane);
                        const uint32_t rn = lane < 31 ? __ldg(rel + ((size_t)slab * row_blocks + rb) * 32 + lane + 1) : 0u;
                        mg0[j] = tb + r;
                        mg1[j] = lane < 31 ? tb + rn : te;
                    } else {
                        const size_t o = (size_t)slab * ((size_t)M + 1) + row;
                        mg0[j] = __ldg(off + o);
                        mg1[j] = __ldg(off + o + 1);
                    }
                }
            }
        };
        // Pieces of at least one block per warp: warp w owns blocks w, w + 8, ... outright (no redundant
        // metadata loads, statistically balanced).  Shorter pieces: every warp scans every block and keeps the
        // active rows whose rank is its own modulo the warp count (exact balance).
        const bool by_block = BB && (blk_b - blk_a) >= kRsWarps;
        const int bstep = by_block ? kRsWarps : 1;
        if (kRsEarly && piece == 0) {                     // x (and the partial rows) may belong to the previous kernel: wait first
            load_meta(blk_a + (by_block ? warp : 0), true, false);
            pdl_wait();
            SPMV_STAMP(wg, 1);
            load_meta(blk_a + (by_block ? warp : 0), false, true);
        } else load_meta(blk_a + (by_block ? warp : 0));
        for (int blk0 = blk_a + (by_block ? warp : 0); blk0 < blk_b; blk0 += kRsBatch * bstep) {
            // ---- this warp's share of the active rows ----------------------------------------------------
            __syncwarp();                                 // the previous list has been issued completely
            int n_rows = 0;
#pragma unroll
            for (int j = 0; j < kRsBatch; j++) {
                const int blk = blk0 + j * bstep, row = blk * 32 + lane;
                const bool valid = blk < blk_b && row >= row_a && row < row_b && mx[j] != 0.0f && mg1[j] > mg0[j];
                const unsigned mask = __ballot_sync(kFull, valid);
                const int rank = rank_base + __popc(mask & lt);
                const bool mine = valid && (by_block || (rank & (kRsWarps - 1)) == warp);
                const unsigned mm = __ballot_sync(kFull, mine);
                if (mine) list[n_rows + __popc(mm & lt)] = make_uint4(mg0[j], mg1[j], __float_as_uint(mx[j]), 0u);
                n_rows += __popc(mm);
                rank_base += __popc(mask);
            }
            if (blk0 + kRsBatch * bstep < blk_b) load_meta(blk0 + kRsBatch * bstep);
            __syncwarp();
            if (piece == 0 && blk0 == blk_a + (by_block ? warp : 0)) SPMV_STAMP(wg, 2);
            // ---- the rows' chunks through the register pipeline (it keeps running across refills) -------
            int r = 0;
            uint32_t g = 0u, g1 = 0u;
            float xr = 0.0f;
            if (n_rows) { const uint4 m = list[0]; g = m.x; g1 = m.y; xr = __uint_as_float(m.z); }
            while (r < n_rows) {
#pragma unroll
                for (int k = 0; k < kRsDepth; k++) {
                    consume(sv[k], si[k], sx[k], sn[k]);
                    if (r < n_rows) {                     // (warp-uniform) issue the next chunk into the slot just retired
                        const uint32_t gs = g + lane;
                        const bool ok = gs < g1;
                        sv[k] = rs_load_vals(vals + gs, ok);
                        si[k] = IX::load(idx, gs, ok);
                        sx[k] = xr;
                        sn[k] = (int)min(32u, g1 - g);
                        g += 32u;
                        if (g >= g1) {
                            r++;
                            if (r < n_rows) { const uint4 m = list[r]; g = m.x; g1 = m.y; xr = __uint_as_float(m.z); }
                        }
                    } else sn[k] = 0;
                }
                if (piece == 0 && r <= kRsDepth + 1 && blk0 == blk_a + (by_block ? warp : 0)) SPMV_STAMP(wg, 3);   // (roughly: the first slots have landed)
            }
        }
        SPMV_STAMP(wg, 4);
        if (kRsTrigger && u >= u_end) pdl_trigger();        // the reduce kernel's CTAs may take their places (they wait for this grid to finish)
        // ---- end of the piece: drain, then the fixed-order sum of the warps -> one partial row ------------
#pragma unroll
        for (int k = 0; k < kRsDepth; k++) { consume(sv[k], si[k], sx[k], sn[k]); sn[k] = 0; }
        SPMV_STAMP(wg, 5);
        __syncthreads();
        SPMV_STAMP(wg, 6);
        const int n_valid = min(W, N - slab * W);
        const int wstride = rs_warp_bytes(W, BB) / 4;
        const float *acc0 = reinterpret_cast<const float *>(smem_raw);
        float *dst = partial + ((size_t)blockIdx.x * kmax + piece) * W;
        for (int c = tid * 4; c < W; c += kRsThreads * 4) {       // (W is a multiple of 4; the warps' rows are 16-byte aligned)
            float4 s = *reinterpret_cast<const float4 *>(acc0 + c);
#pragma unroll
            for (int w = 1; w < kRsWarps; w++) s = f4_add(s, *reinterpret_cast<const float4 *>(acc0 + (size_t)w * wstride + c));
            if (c + 3 >= n_valid) {                                 // columns past N hold whatever the pads added: zero them
                if (c >= n_valid) s.x = 0.0f;
                if (c + 1 >= n_valid) s.y = 0.0f;
                if (c + 2 >= n_valid) s.z = 0.0f;
                s.w = 0.0f;
            }
            *reinterpret_cast<float4 *>(dst + c) = s;
        }
        SPMV_STAMP(wg, 7);
        __syncthreads();                                  // the accumulators are zeroed again by the next piece
    }
}

// y[slab columns] = sum of the slab's partial rows in CTA order.  A CTA handles 32 float4 columns of one
// slab with 8 row groups: group k adds rows k, k+8, ... (8 loads in flight), then the 8 group sums are
// added in group order: the order depends only on (shape, G).
__global__ void __launch_bounds__(256)
panel_rs_reduce_kernel(const float *__restrict__ partial, const YDst yd, int M, int N, int W, int slabs, int kmax, long long G, int small)
{
    __shared__ float4 sums[256];
    const int wg = 32768 + blockIdx.x * 8 + (threadIdx.x >> 5);   // (timeline builds only)
    SPMV_STAMP(wg, 0);
    if (kRsEarly) pdl_trigger();
    pdl_wait();
    SPMV_STAMP(wg, 1);
    const int v = threadIdx.x & 31, k = threadIdx.x >> 5;
    const int per_slab = W / 128;                         // CTAs per slab (W is a power of two >= 256)
    const int slab = blockIdx.x / per_slab, col4 = (blockIdx.x - slab * per_slab) * 32 + v;
    const long long T = (long long)slabs * M;
    const long long c_lo = rs_owner((long long)slab * M, T, G, small), c_hi = rs_owner((long long)slab * M + M - 1, T, G, small);
    const int rows = (int)(c_hi - c_lo + 1);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = k; j < rows; j += 8 * kRedBatch) {
        float4 t[kRedBatch];
#pragma unroll
        for (int q = 0; q < kRedBatch; q++) {
            const int jq = j + q * 8;
            t[q] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (jq < rows) {
                const long long c = c_lo + jq;
                const long long cb = rs_begin(c, T, G, small);
                const size_t row = (size_t)c * kmax + (size_t)(slab - (small ? (int)((unsigned)cb / (unsigned)M) : (int)(cb / M)));
                t[q] = __ldcg(reinterpret_cast<const float4 *>(partial + row * W) + col4);
            }
        }
#pragma unroll
        for (int q = 0; q < kRedBatch; q++) acc = f4_add(acc, t[q]);
    }
    sums[threadIdx.x] = acc;
    __syncthreads();
    if (k == 0 && (long long)slab * W + col4 * 4 < N) {
        float4 s = sums[v];
#pragma unroll
        for (int g = 1; g < 8; g++) s = f4_add(s, sums[g * 32 + v]);
        y_store4(yd, ((size_t)slab * W >> 2) + col4, s);
    }
    SPMV_STAMP(wg, 2);
}

template <int IDXB, bool TILED, bool BB>
int launch_rs(spmv_plan *p, const float *x, const YDst &y, cudaStream_t st)
{
    const DevPanel &d = p->panel;
    const int small = ((long long)d.slabs * p->M + 1) * ((long long)d.rs_grid + 1) < (1ll << 31) ? 1 : 0;
    auto k = panel_rs_kernel<IDXB, TILED, BB>;
    static int smem_set[16] = {0};
    if (d.rs_smem > 48 * 1024 && p->device >= 0 && p->device < 16 && smem_set[p->device] < d.rs_smem) {
        SPMV_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, d.rs_smem));
        smem_set[p->device] = d.rs_smem;
    }
    SPMV_CUDA(launch_k(k, dim3((unsigned)d.rs_grid), dim3(kRsThreads), (size_t)d.rs_smem, st, reinterpret_cast<const float4 *>(d.vals),
                       (const void *)d.idx, (const uint32_t *)d.off, (const uint16_t *)d.rel, x, d.rs_partial, (int)p->M, (int)p->N,
                       d.slab_cols, d.row_blocks, d.slabs, d.rs_kmax, small));
    SPMV_CUDA(launch_k(panel_rs_reduce_kernel, dim3((unsigned)(d.slabs * (d.slab_cols / 128))), dim3(256), 0, st,
                       (const float *)d.rs_partial, y, (int)p->M, (int)p->N, d.slab_cols, d.slabs, d.rs_kmax, (long long)d.rs_grid, small));
    return SPMV_OK;
}

} // namespace

int launch_panel_rs(spmv_plan *p, const float *d_x, const YDst &yd, cudaStream_t st)
{
    const DevPanel &d = p->panel;
    if (d.rs_by_block) {
        if (d.index_bits == 8) return d.tiled ? launch_rs<8, true, true>(p, d_x, yd, st) : launch_rs<8, false, true>(p, d_x, yd, st);
        return d.tiled ? launch_rs<16, true, true>(p, d_x, yd, st) : launch_rs<16, false, true>(p, d_x, yd, st);
    }
    if (d.index_bits == 8) return d.tiled ? launch_rs<8, true, false>(p, d_x, yd, st) : launch_rs<8, false, false>(p, d_x, yd, st);
    return d.tiled ? launch_rs<16, true, false>(p, d_x, yd, st) : launch_rs<16, false, false>(p, d_x, yd, st);
}

// Geometry of the register-staged form: G = kRsCtas CTAs per SM (as many as shared memory allows), rounded to a
// whole number of CTAs per slab when that costs under 10 % of the grid; at least 32 rows per CTA.
int configure_panel_rs(spmv_plan *p, const HostPanel &h)
{
    DevPanel &d = p->panel;
    const int smem_cap = p->max_smem_optin > 0 ? p->max_smem_optin : 227 * 1024;
    const int64_t slabs = std::max(1, h.slabs), M = std::max<int64_t>(1, h.M), T = slabs * M;
    d.rs_grid = 0;
    for (int pass = 0; pass < 2; pass++) {                // the list size depends on the piece length, which depends on the grid
        d.rs_smem = kRsWarps * rs_warp_bytes(h.slab_cols, d.rs_by_block);
        if (d.rs_smem > smem_cap) return SPMV_OK;         // the ring kernel takes it
        const int resident = std::max(1, std::min(kRsCtas, (228 * 1024) / (d.rs_smem + 1024)));
        int64_t G = (int64_t)p->sm_count * resident;
        if (G >= slabs && (G / slabs) * slabs * 10 >= G * 9) G = (G / slabs) * slabs;
        G = std::max<int64_t>(1, std::min<int64_t>(G, (T + 31) / 32));
        const bool bb = T / G >= 32 * kRsWarps;
        d.rs_grid = (int)G;
        d.rs_kmax = (int)(2 + ((T + G - 1) / G) / M);
        if (bb == d.rs_by_block) break;
        d.rs_by_block = bb;
    }
    d.rs_smem = kRsWarps * rs_warp_bytes(h.slab_cols, d.rs_by_block);
    if (d.rs_smem > smem_cap) { d.rs_grid = 0; return SPMV_OK; }
    const size_t floats = (size_t)d.rs_grid * d.rs_kmax * h.slab_cols;
    int rc = plan_alloc(p, reinterpret_cast<void **>(&d.rs_partial), floats * sizeof(float), false);
    if (!rc) p->scratch_bytes += (int64_t)(floats * sizeof(float));
    return rc;
}

} // namespace spmv
