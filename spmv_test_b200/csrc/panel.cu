// panel.cu — activation+weight-sparse SGEMV on the row-panel format (formats.hpp: HostPanel).
//
// Replaces awsp_kernel_v0/v1/v2 (reference awsp.cu:5-317), awsp_ref_kernel (awsp_ref.cu:6-185)
// and csr_tiling_kernel (csr_tiling.cu:24-114).  The reference gives every 32-column slab to
// one CTA, walks a per-row bitmap (word -> popc -> address -> one 4-byte load per lane) and
// uses x only as a load predicate: every bitmap word and every x is still read.  Here
//   * a CTA owns (column slab, row range); each of its warps walks whole 32-row blocks;
//   * lane l of a warp looks at row l of the block: x[row] and the segment's group range.
//     ballot(x != 0 && segment non-empty) is the activation compaction: rows with x == 0 are
//     never visited, so their values/indices are never read from HBM;
//   * a visited segment is streamed as 128-bit groups (float4 values + 4 packed column ids),
//     32 groups per chunk, kStages chunks in flight per warp through a cp.async ring in shared
//     memory (commit/wait groups give a true FIFO; a register ring collapses to one load in
//     flight because its loads share scoreboard slots — measured, profiles/r01_notes.md);
//   * products are accumulated into a per-warp fp32 accumulator row in shared memory
//     (columns inside one segment are distinct, segments are consumed in ascending row order,
//     warps never share an accumulator) — no atomics, fixed summation order;
//   * warps are summed in warp order, row splits in split order (integer ticket picks the CTA
//     that does the final sum; the order of the sum itself is fixed).
// AWSP addresses segments through a 32-bit per-row table, TCSR through 32-bit per-tile plus
// 16-bit in-tile offsets (the reference's blk_idx, tcsr.cpp:13,34, made two-level).
#include <algorithm>
#include <cmath>

#include "common.cuh"
#include "plan.hpp"

namespace spmv {

namespace {

constexpr int kPanelMaxWarps = 8;
constexpr int kPanelThreads = kPanelMaxWarps * 32;   // launch bound; the CTA size is a plan parameter
constexpr int kStages = 8;                   // chunks in flight per warp (cp.async groups)

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

// per-warp shared memory: [acc: W floats][vals ring: kStages x 32 float4][idx ring][meta: 32 x uint4][slot info]
template <int IDXB> __host__ __device__ constexpr int warp_smem_bytes(int W)
{
    return W * 4 + kStages * 32 * 16 + kStages * 32 * (IDXB == 8 ? 4 : 8) + 32 * 16 + kStages * 8;
}

template <int IDXB, bool TILED>
__global__ void __launch_bounds__(kPanelThreads)
panel_kernel(const float4 *__restrict__ vals, const void *__restrict__ idx,
             const uint32_t *__restrict__ off, const uint16_t *__restrict__ rel,
             const float *__restrict__ x, float *__restrict__ y, float *__restrict__ partial,
             unsigned *__restrict__ tickets, int M, int N, int W, int row_blocks,
             int blocks_per_split, int splits)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int last_flag;
    using CI = ColIdx<IDXB>;
    using IVec = typename CI::Vec;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_warps = blockDim.x >> 5;
    const int slab = blockIdx.x, split = blockIdx.y;
    unsigned char *wbase = smem_raw + (size_t)warp * warp_smem_bytes<IDXB>(W);
    float *acc = reinterpret_cast<float *>(wbase);
    float4 *ring_v = reinterpret_cast<float4 *>(wbase + (size_t)W * 4);
    IVec *ring_i = reinterpret_cast<IVec *>(wbase + (size_t)W * 4 + kStages * 32 * 16);
    uint4 *meta = reinterpret_cast<uint4 *>(wbase + (size_t)W * 4 + kStages * 32 * 16 + kStages * 32 * sizeof(IVec));
    uint2 *sinfo = reinterpret_cast<uint2 *>(meta + 32);   // per ring slot: (valid lanes, x of the row)
    const int wg = (blockIdx.y * gridDim.x + blockIdx.x) * n_warps + warp;   // trace id
    (void)wg;
    SPMV_STAMP(wg, 0);
    for (int c = lane; c < W; c += 32) acc[c] = 0.0f;
    __syncwarp();

    const int rb_end = min(row_blocks, (split + 1) * blocks_per_split);
    const int rb_first = split * blocks_per_split + warp; // this warp's blocks: rb_first, +n_warps, ...

    // ---- metadata of one 32-row block: lane = row ------------------------------------------
    // Two blocks of metadata live in registers: A (next to become current) and B (the one
    // after), so a block's x / offset loads are issued two blocks before they are needed.
    // (Requesting the segments' lines into L2 ahead of the ring with prefetch.global.L2 was
    // tried and is slower: every line is one more L1TEX request on an LSU-bound kernel.)
    struct Meta { float xv; uint32_t g0, g1; };
    auto load_meta = [&](int rb) {
        Meta m;
        const int row = rb * 32 + lane;
        m.xv = row < M ? __ldg(x + row) : 0.0f;
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
    Meta mA = {0.f, 0u, 0u}, mB = {0.f, 0u, 0u};
    if (rb_first < rb_end) mA = load_meta(rb_first);
    if (rb_first + n_warps < rb_end) mB = load_meta(rb_first + n_warps);

    // ---- kStages chunks in flight: cp.async ring, one commit group per chunk -------------------
    // The (block, active row, chunk of 32 groups) nest below is one flat sequence of chunks to
    // the ring: `step` first retires the oldest slot (read-modify-write of the accumulator row),
    // then refills it with the chunk at hand.  Branch-free per chunk: lanes past the segment's
    // end zero-fill their slots, compute like everyone and only their stores are predicated off,
    // so an idle lane can never overwrite a live lane's update.  Slot info (valid lanes, x of
    // the row) sits next to the ring; everything starts zeroed, so the warm-up steps and the
    // final drain are the same code.
    for (int k = lane; k < kStages * 32; k += 32) ring_i[k] = CI::zero();
    if (lane < kStages) sinfo[lane] = make_uint2(0u, 0u);
    __syncwarp();
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
        __syncwarp();                                     // next chunk may be another row
    };

    bool stamped = false; (void)stamped;
    for (int rb = split * blocks_per_split + warp; rb < rb_end; rb += n_warps) {
        // activation compaction of this block: ballot over x != 0 (and a non-empty segment),
        // then an order-preserving popc scatter of (first group, end group, x) into the list
        const bool active = mA.xv != 0.0f && mA.g1 > mA.g0;
        const unsigned mask = __ballot_sync(kFull, active);
        if (active) meta[__popc(mask & ((1u << lane) - 1u))] = make_uint4(mA.g0, mA.g1, __float_as_uint(mA.xv), 0u);
        __syncwarp();
        const int n_rows = __popc(mask);
        mA = mB;                                          // issued two blocks ago
        if (rb + 2 * n_warps < rb_end) mB = load_meta(rb + 2 * n_warps);
#ifdef SPMV_TRACE
        if (!stamped) { SPMV_STAMP(wg, 1); stamped = true; }
#endif
        for (int r = 0; r < n_rows; r++) {
            const uint4 m = meta[r];
            const float xv = __uint_as_float(m.z);
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
                if (lane == 0) sinfo[s] = make_uint2(min(32u, m.y - g), m.z);
            }
        }
        __syncwarp();                                     // the list is rewritten next
    }
    SPMV_STAMP(wg, 2);
    cp_async_wait<0>();
    SPMV_STAMP(wg, 3);
#pragma unroll 1
    for (int k = 0; k < kStages; k++) retire(it++ & (kStages - 1));
    SPMV_STAMP(wg, 4);

    // ---- fixed-order sum over warps, then over row splits ----------------------------------------
    __syncthreads();
    SPMV_STAMP(wg, 5);
    const int col0 = slab * W;
    const int n_valid = min(W, N - col0);
    const size_t npad = (size_t)gridDim.x * W;
    const int wstride = warp_smem_bytes<IDXB>(W) / 4;
    const float *acc0 = reinterpret_cast<const float *>(smem_raw);
    for (int c = tid; c < n_valid; c += blockDim.x) {
        float s = acc0[c];
        for (int w = 1; w < n_warps; w++) s += acc0[(size_t)w * wstride + c];
        if (splits == 1) y[col0 + c] = s;
        else partial[(size_t)split * npad + col0 + c] = s;
    }
    SPMV_STAMP(wg, 6);
    if (splits > 1)
        split_reduce_finish(y, partial, tickets, slab, splits, W, n_valid, npad, &last_flag,
                            reinterpret_cast<float4 *>(smem_raw));   // accumulators are dead by now
    SPMV_STAMP(wg, 7);
}

template <int IDXB, bool TILED>
int launch_variant(spmv_plan *p, const float *x, float *y, cudaStream_t st)
{
    auto k = panel_kernel<IDXB, TILED>;
    if (p->smem > 48 * 1024)
        SPMV_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, p->smem));
    const DevPanel &d = p->panel;
    k<<<p->grid, p->block, p->smem, st>>>(reinterpret_cast<const float4 *>(d.vals), d.idx, d.off, d.rel, x, y,
                                              p->partial, p->tickets, (int)p->M, (int)p->N, d.slab_cols,
                                              d.row_blocks, d.blocks_per_split, p->row_splits);
    SPMV_CUDA(cudaGetLastError());
    return SPMV_OK;
}

} // namespace

int launch_panel(spmv_plan *p, const float *d_x, float *d_y, cudaStream_t st)
{
    if (p->N == 0) return SPMV_OK;
    const DevPanel &d = p->panel;
    if (d.index_bits == 8) return d.tiled ? launch_variant<8, true>(p, d_x, d_y, st) : launch_variant<8, false>(p, d_x, d_y, st);
    return d.tiled ? launch_variant<16, true>(p, d_x, d_y, st) : launch_variant<16, false>(p, d_x, d_y, st);
}

// Geometry: grid = (slabs, row splits).  A warp should own at least two 32-row blocks (so
// the metadata prefetch has something to overlap), the grid should cover the SMs about
// twice, and a CTA should stream clearly more than it spends zeroing / summing its
// accumulator rows (kPanelWarps * slab_cols floats).
int configure_panel(spmv_plan *p, const HostPanel &h, const spmv_options_t *o)
{
    DevPanel &d = p->panel;
    d.slab_cols = h.slab_cols; d.index_bits = h.index_bits; d.slabs = h.slabs;
    d.row_blocks = h.row_blocks; d.tiled = h.tiled;
    // Geometry.  Every CTA pays a few microseconds of serial latency (metadata, first chunks,
    // cross-warp and cross-split sums), so the grid is kept to one resident wave and the search
    // below maximises how evenly that wave loads the SMs, the row blocks and the warps, at about
    // 24 warps per SM in total (measured on B200, profiles/r01_notes.md).
    const int rb = std::max(1, h.row_blocks);
    const int slabs = std::max(1, h.slabs);
    const int smem_cap = p->max_smem_optin > 0 ? p->max_smem_optin : 227 * 1024;
    const int per_warp = h.index_bits == 8 ? warp_smem_bytes<8>(h.slab_cols) : warp_smem_bytes<16>(h.slab_cols);
    int warps = 0, splits = 0;
    double best = -1.0;
    for (int w : {4, 8}) {
        if (o && o->warps_per_col > 0 && w != std::min(kPanelMaxWarps, std::max(1, o->warps_per_col))) continue;
        if (w * per_warp > smem_cap) continue;
        const int resident = std::max(1, std::min(2048 / (w * 32), (228 * 1024) / (w * per_warp + 1024)));
        for (int s0 = 1; s0 <= rb; s0++) {
            if (o && o->row_splits > 0 && s0 != std::min(o->row_splits, rb)) continue;
            const int bps = (rb + s0 - 1) / s0, s1 = (rb + bps - 1) / bps;
            if (s1 != s0) continue;                                   // canonical split counts only
            const int64_t ctas = (int64_t)slabs * s1;
            const double per_sm = (double)ctas / p->sm_count;
            const double sm_bal = per_sm / std::ceil(per_sm);                       // SMs equally loaded
            const double row_bal = (double)rb / ((double)bps * s1);                  // splits equally long
            const double warp_bal = (double)bps / (std::ceil((double)bps / w) * w);  // warps equally loaded
            const double tw = (double)ctas * w / p->sm_count;                        // warps per SM
            const double fill = std::min(1.0, tw / 24.0);
            const double waves = (double)ctas / ((double)p->sm_count * resident);
            const double wave_pen = waves <= 1.0 ? 1.0 : 1.0 / (0.6 + 0.4 * std::ceil(waves));
            const double latency_pen = 1.0 / (1.0 + 0.02 * std::max(0.0, tw - 24.0));
            const double score = sm_bal * row_bal * warp_bal * fill * wave_pen * latency_pen;
            if (score > best + 1e-9) { best = score; warps = w; splits = s1; }
        }
    }
    if (warps == 0) {                                                 // forced options outside the search space
        warps = (o && o->warps_per_col > 0) ? std::min(kPanelMaxWarps, std::max(1, o->warps_per_col)) : 4;
        splits = (o && o->row_splits > 0) ? std::min(o->row_splits, rb) : 1;
        if (warps * per_warp > smem_cap)
            return set_error(SPMV_ERR_UNSUPPORTED, "panel: %d bytes of shared memory exceed the device limit", warps * per_warp);
    }
    d.warps = warps;
    p->block = warps * 32;
    p->smem = warps * per_warp;
    p->tile_width = h.slab_cols;
    p->col_tiles = h.slabs;
    p->kernels_per_run = 1;
    d.blocks_per_split = (rb + splits - 1) / splits;
    splits = (rb + d.blocks_per_split - 1) / d.blocks_per_split;      // no empty splits
    p->row_splits = splits;
    p->grid = dim3((unsigned)slabs, (unsigned)splits, 1);
    return alloc_split_scratch(p);
}

} // namespace spmv
