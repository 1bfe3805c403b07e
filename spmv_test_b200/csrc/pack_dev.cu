// pack_dev.cu — device packers: a dense row-major matrix that already sits in
// HBM is turned into this library's device formats by kernels, without a host round trip.
//
// The reference packs on one host thread with vector<vector<float>> copies (wsp.cpp:25-37,
// awsp.cpp:30-46, tcsr.cpp:20-36, matrix_csr.cpp:8-22) and that is what every one of its
// launcher calls spends its wall time on.  Here the same three steps every packer has —
// count, prefix-sum, fill — are kernels:
//   count   one warp per row segment (panel formats) / one thread per (row range, column) (wsp)
//   scan    three-pass exclusive prefix sum of the group counts -> the offset tables
//   fill    the same work split writes values + indices at the scanned positions
//   deal    the bank-aware order inside every chunk of 32 groups, one thread per chunk, by the very
//           routine the host packers use (deal.hpp)
// so the arrays are BIT-IDENTICAL to the host packers' (tests/test_gpu_parity.py compares plan
// files byte for byte).  Integer work only; nothing here is on the SGEMV hot path.
#include <algorithm>

#include "common.cuh"
#include "deal.hpp"
#include "plan.hpp"

namespace spmv {

namespace {

// ------------------------------------------------------------------------------------------
// exclusive prefix sum of 32-bit counts, 64-bit total
// ------------------------------------------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ unsigned long long block_sum_u64(unsigned long long v, unsigned long long *ws)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(kFull, v, d);
    if (lane == 0) ws[warp] = v;
    __syncthreads();
    unsigned long long t = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += ws[w];
    __syncthreads();
    return t;
}

__global__ void __launch_bounds__(kScanThreads) scan_tile_sums(const uint32_t *__restrict__ in, size_t n,
                                                                unsigned long long *__restrict__ tile_sum)
{
    __shared__ unsigned long long ws[kScanThreads / 32];
    const size_t base = (size_t)blockIdx.x * kScanTile + (size_t)threadIdx.x * kScanItems;
    unsigned long long s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; k++)
        if (base + k < n) s += in[base + k];
    const unsigned long long t = block_sum_u64(s, ws);
    if (threadIdx.x == 0) tile_sum[blockIdx.x] = t;
}

// one CTA: tile sums -> exclusive tile offsets (in place) + grand total
__global__ void __launch_bounds__(1024) scan_tiles(unsigned long long *__restrict__ tile, size_t tiles,
                                                   unsigned long long *__restrict__ total)
{
    __shared__ unsigned long long part[1024];
    const size_t per = (tiles + blockDim.x - 1) / blockDim.x;
    const size_t a0 = (size_t)threadIdx.x * per, a = a0 < tiles ? a0 : tiles, b = a + per < tiles ? a + per : tiles;
    unsigned long long s = 0;
    for (size_t i = a; i < b; i++) s += tile[i];
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long run = 0;
        for (int t = 0; t < (int)blockDim.x; t++) { const unsigned long long v = part[t]; part[t] = run; run += v; }
        *total = run;
    }
    __syncthreads();
    unsigned long long run = part[threadIdx.x];
    for (size_t i = a; i < b; i++) { const unsigned long long v = tile[i]; tile[i] = run; run += v; }
}

__global__ void __launch_bounds__(kScanThreads) scan_apply(const uint32_t *in, uint32_t *out, size_t n,
                                                            const unsigned long long *__restrict__ tile_off)
{
    __shared__ unsigned long long part[kScanThreads];
    const size_t base = (size_t)blockIdx.x * kScanTile + (size_t)threadIdx.x * kScanItems;
    uint32_t v[kScanItems];
    unsigned long long s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; k++) { v[k] = base + k < n ? in[base + k] : 0u; s += v[k]; }
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x < 32) {                               // 256 partial sums: 8 per lane, then a warp scan
        unsigned long long loc[8], t = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) { loc[k] = t; t += part[threadIdx.x * 8 + k]; }
        unsigned long long incl = t;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long o = __shfl_up_sync(kFull, incl, d);
            if ((int)threadIdx.x >= d) incl += o;
        }
        const unsigned long long excl = incl - t;
#pragma unroll
        for (int k = 0; k < 8; k++) part[threadIdx.x * 8 + k] = excl + loc[k];
    }
    __syncthreads();
    unsigned long long run = tile_off[blockIdx.x] + part[threadIdx.x];
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
        if (base + k < n) out[base + k] = (uint32_t)run;
        run += v[k];
    }
}

struct DevTmp {                                           // frees its buffers on scope exit
    void *ptr[12]; int n = 0;
    ~DevTmp() { for (int i = 0; i < n; i++) cudaFree(ptr[i]); cudaGetLastError(); }
    template <class T> int get(T **out, size_t count, bool zero)
    {
        const size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
        SPMV_CUDA(cudaMalloc(reinterpret_cast<void **>(out), bytes));
        ptr[n++] = *out;
        if (zero) SPMV_CUDA(cudaMemset(*out, 0, bytes));
        return SPMV_OK;
    }
};

// out may alias in.  *total_h receives the sum of all counts.
int exclusive_scan_u32(const uint32_t *in, uint32_t *out, size_t n, unsigned long long *total_h)
{
    DevTmp tmp;
    const size_t tiles = (n + kScanTile - 1) / kScanTile;
    unsigned long long *tile = nullptr, *total = nullptr;
    int rc = tmp.get(&tile, tiles, false);
    if (!rc) rc = tmp.get(&total, 1, true);
    if (rc) return rc;
    if (tiles) {
        scan_tile_sums<<<(unsigned)tiles, kScanThreads>>>(in, n, tile);
        scan_tiles<<<1, 1024>>>(tile, tiles, total);
        scan_apply<<<(unsigned)tiles, kScanThreads>>>(in, out, n, tile);
        SPMV_CUDA(cudaGetLastError());
    }
    SPMV_CUDA(cudaMemcpy(total_h, total, sizeof *total_h, cudaMemcpyDeviceToHost));
    return SPMV_OK;
}

// ------------------------------------------------------------------------------------------
// non-zero count of the dense matrix (a14's test: val != 0.0f, so -0.0 is a zero and NaN is kept)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) count_nnz_kernel(const float *__restrict__ A, long long lda, int M, int N,
                                                        unsigned long long *__restrict__ total)
{
    __shared__ unsigned long long ws[8];
    unsigned long long n = 0;
    for (int r = blockIdx.y; r < M; r += gridDim.y)
        for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < N; c += gridDim.x * blockDim.x)
            n += A[(long long)r * lda + c] != 0.0f;
    n = block_sum_u64(n, ws);
    if (threadIdx.x == 0 && n) atomicAdd(total, n);     // one atomic per CTA
}

int count_nnz(const float *d_A, int64_t lda, int64_t M, int64_t N, int64_t *nnz)
{
    DevTmp tmp;
    unsigned long long *total = nullptr, h = 0;
    int rc = tmp.get(&total, 1, true);
    if (rc) return rc;
    if (M > 0 && N > 0) {
        dim3 grid((unsigned)std::min<int64_t>((N + 255) / 256, 16), (unsigned)std::min<int64_t>(M, 296));
        count_nnz_kernel<<<grid, 256>>>(d_A, lda, (int)M, (int)N, total);
        SPMV_CUDA(cudaGetLastError());
    }
    SPMV_CUDA(cudaMemcpy(&h, total, sizeof h, cudaMemcpyDeviceToHost));
    *nnz = (int64_t)h;
    return SPMV_OK;
}

// ------------------------------------------------------------------------------------------
// panel formats (AWSP / TCSR): one warp per row segment
// ------------------------------------------------------------------------------------------
constexpr int kPackWarps = 8;

__global__ void __launch_bounds__(kPackWarps * 32)
panel_count_kernel(const float *__restrict__ A, long long lda, int M, int N, int W, int slabs,
                   uint32_t *__restrict__ cntg, int *__restrict__ row_nnz, int *__restrict__ row_groups,
                   int *__restrict__ row_segs)
{
    const int lane = threadIdx.x & 31;
    const long long wid = (long long)blockIdx.x * kPackWarps + (threadIdx.x >> 5);
    if (wid >= (long long)slabs * M) return;
    const int s = (int)(wid / M), r = (int)(wid - (long long)s * M);
    const int c0 = s * W, cw = min(W, N - c0);
    const float *row = A + (long long)r * lda + c0;
    int n = 0;
    for (int c = lane; c < cw; c += 32) n += row[c] != 0.0f;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) n += __shfl_xor_sync(kFull, n, d);
    if (lane == 0) {
        const int g = (n + 3) >> 2;
        cntg[(size_t)s * ((size_t)M + 1) + r] = (uint32_t)g;
        if (n) { atomicAdd(row_nnz + r, n); atomicAdd(row_groups + r, g); atomicAdd(row_segs + r, 1); }
    }
}

template <class IdxT>
__global__ void __launch_bounds__(kPackWarps * 32)
panel_fill_kernel(const float *__restrict__ A, long long lda, int M, int N, int W, int slabs, int warps,
                  const uint32_t *__restrict__ off, float *__restrict__ vals, IdxT *__restrict__ idx)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    const long long wid = (long long)blockIdx.x * warps + warp;
    if (wid >= (long long)slabs * M) return;
    const int s = (int)(wid / M), r = (int)(wid - (long long)s * M);
    const size_t o = (size_t)s * ((size_t)M + 1) + r;
    const uint32_t g0 = off[o], g1 = off[o + 1];
    if (g1 == g0) return;                                 // empty segment
    // per warp: (W + 4) values, then (W + 4) 16-bit columns
    const size_t per_warp = ((size_t)(W + 4) * 6 + 15) & ~(size_t)15;
    float *vs = reinterpret_cast<float *>(smem_raw + (size_t)warp * per_warp);
    uint16_t *cs = reinterpret_cast<uint16_t *>(vs + (W + 4));

    const int c0 = s * W, cw = min(W, N - c0);
    const float *row = A + (long long)r * lda + c0;
    int n = 0, absent = -1;
    for (int c = 0; c < cw; c += 32) {                    // ballot compaction, ascending columns
        const int col = c + lane;
        const float v = col < cw ? row[col] : 0.0f;
        const bool nz = v != 0.0f;
        const unsigned mask = __ballot_sync(kFull, nz);
        if (nz) { const int p = n + __popc(mask & lt); vs[p] = v; cs[p] = (uint16_t)col; }
        const unsigned zeros = ~mask & __ballot_sync(kFull, col < cw);
        if (absent < 0 && zeros) absent = c + __ffs(zeros) - 1;   // smallest column absent from the segment
        n += __popc(mask);
    }
    const int g = (int)(g1 - g0);
    if (n + lane < 4 * g) { vs[n + lane] = 0.0f; cs[n + lane] = (uint16_t)absent; }   // at most 3 pads
    __syncwarp();
    for (int ch = lane; ch * 32 < g; ch += 32) {          // one lane per chunk of 32 groups
        const int lanes = min(32, g - ch * 32);
        const float *v_in = vs + ch * 128;
        const uint16_t *c_in = cs + ch * 128;
        float *v_out = vals + ((size_t)g0 + (size_t)ch * 32) * 4;
        IdxT *i_out = idx + ((size_t)g0 + (size_t)ch * 32) * 4;
        deal_chunk(lanes, [&](int k) { return (unsigned)c_in[k]; },
                   [&](int slot, int k) { v_out[slot] = v_in[k]; i_out[slot] = (IdxT)c_in[k]; });
    }
}

// TCSR's two-level offsets from the per-row table: tile_off[slab][rb] and rel[slab][rb][r];
// flag[0] is raised when a tile holds more than 65535 groups (16-bit rel would overflow)
__global__ void __launch_bounds__(256)
panel_tile_offsets_kernel(const uint32_t *__restrict__ off_row, int M, int row_blocks, int slabs,
                          uint32_t *__restrict__ tile_off, uint16_t *__restrict__ rel, int *__restrict__ flag)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)slabs * row_blocks * kTileRows) return;
    const int r = (int)(t % kTileRows);
    const long long tb = t / kTileRows;
    const int rb = (int)(tb % row_blocks), s = (int)(tb / row_blocks);
    const uint32_t *o = off_row + (size_t)s * ((size_t)M + 1);
    const uint32_t base = o[(size_t)rb * kTileRows];
    const uint32_t cur = o[min((long long)M, (long long)rb * kTileRows + r)];
    rel[t] = (uint16_t)(cur - base);
    if (r == 0) {
        tile_off[(size_t)s * (row_blocks + 1) + rb] = base;
        const uint32_t end = o[min((long long)M, (long long)(rb + 1) * kTileRows)];
        if (end - base > 65535u) atomicExch(flag, 1);
        if (rb == row_blocks - 1) tile_off[(size_t)s * (row_blocks + 1) + row_blocks] = o[M];
    }
}

// ------------------------------------------------------------------------------------------
// WSP (CSR of A^T in groups of four, optional row panels): one thread per (row range, column)
// ------------------------------------------------------------------------------------------
constexpr int kWspSub = 128;                              // rows per counting range

__global__ void __launch_bounds__(256)
wsp_count_kernel(const float *__restrict__ A, long long lda, int M, int N, int nsub, long long panel_rows,
                 uint32_t *__restrict__ cnt_sub)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= N) return;
    const int psub = blockIdx.y, p = psub / nsub, sub = psub - p * nsub;
    const long long r0 = p * panel_rows + (long long)sub * kWspSub;
    const long long r1 = min(r0 + kWspSub, min((long long)M, (p + 1) * panel_rows));
    uint32_t n = 0;
    for (long long r = r0; r < r1; r++) n += A[r * lda + c] != 0.0f;
    cnt_sub[(size_t)psub * N + c] = n;
}

// per list (panel, column): counts of its row ranges -> exclusive positions inside the list (in
// place) and the list's group count
__global__ void __launch_bounds__(256)
wsp_list_kernel(uint32_t *__restrict__ cnt_sub, int N, int nsub, int panels, uint32_t *__restrict__ cntg)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int p = blockIdx.y;
    if (c >= N) return;
    uint32_t run = 0;
    for (int sub = 0; sub < nsub; sub++) {
        uint32_t *q = cnt_sub + ((size_t)p * nsub + sub) * N + c;
        const uint32_t v = *q;
        *q = run;
        run += v;
    }
    cntg[(size_t)p * N + c] = (run + 3u) >> 2;
}

template <class IdxT>
__global__ void __launch_bounds__(256) wsp_init_kernel(float *__restrict__ vals, IdxT *__restrict__ idx, size_t n, IdxT pad)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        vals[i] = 0.0f;
        idx[i] = pad;
    }
}

template <class IdxT>
__global__ void __launch_bounds__(256)
wsp_fill_kernel(const float *__restrict__ A, long long lda, int M, int N, int nsub, long long panel_rows,
                const uint32_t *__restrict__ colptr, const uint32_t *__restrict__ cnt_sub,
                float *__restrict__ vals, IdxT *__restrict__ idx)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= N) return;
    const int psub = blockIdx.y, p = psub / nsub, sub = psub - p * nsub;
    const long long rbase = p * panel_rows;
    const long long r0 = rbase + (long long)sub * kWspSub;
    const long long r1 = min(r0 + kWspSub, min((long long)M, (p + 1) * panel_rows));
    size_t at = (size_t)colptr[(size_t)p * N + c] * 4 + cnt_sub[(size_t)psub * N + c];
    for (long long r = r0; r < r1; r++) {
        const float v = A[r * lda + c];
        if (v != 0.0f) { vals[at] = v; idx[at] = (IdxT)(r - rbase); at++; }
    }
}

// one warp per list, one lane per chunk of 32 groups, in place
template <class IdxT>
__global__ void __launch_bounds__(256)
wsp_deal_kernel(const uint32_t *__restrict__ colptr, long long lists, float *__restrict__ vals, IdxT *__restrict__ idx)
{
    const int lane = threadIdx.x & 31;
    const long long l = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (l >= lists) return;
    const uint32_t g0 = colptr[l], g1 = colptr[l + 1];
    for (uint32_t c0 = g0 + 32u * lane; c0 < g1; c0 += 32u * 32u) {
        const int lanes = (int)min(32u, g1 - c0);
        float *v = vals + (size_t)c0 * 4;
        IdxT *ix = idx + (size_t)c0 * 4;
        float lv[128]; IdxT li[128];
        for (int k = 0; k < 4 * lanes; k++) { lv[k] = v[k]; li[k] = ix[k]; }
        deal_chunk(lanes, [&](int k) { return (unsigned)li[k]; },
                   [&](int slot, int k) { v[slot] = lv[k]; ix[slot] = li[k]; });
    }
}

template <class IdxT>
int wsp_build(spmv_plan *p, const float *d_A, int64_t lda, HostWsp &w, int nsub, const uint32_t *cnt_sub)
{
    const int64_t L = (int64_t)w.panels * w.N;
    const size_t entries = (size_t)(w.groups + 1) * 4;    // one spare pad group (pack_host.cpp: wsp_finish_layout)
    int rc = plan_alloc(p, reinterpret_cast<void **>(&p->wsp.vals), entries * sizeof(float), false);
    if (!rc) rc = plan_alloc(p, &p->wsp.idx, entries * sizeof(IdxT), false);
    if (rc) return rc;
    p->device_bytes += (int64_t)(entries * (sizeof(float) + sizeof(IdxT)));
    IdxT *idx = reinterpret_cast<IdxT *>(p->wsp.idx);
    wsp_init_kernel<IdxT><<<(unsigned)std::min<size_t>((entries + 255) / 256, 148 * 16), 256>>>(p->wsp.vals, idx, entries,
                                                                                               (IdxT)w.panel_rows);
    if (w.M > 0 && w.N > 0) {
        dim3 grid((unsigned)((w.N + 255) / 256), (unsigned)(w.panels * nsub));
        wsp_fill_kernel<IdxT><<<grid, 256>>>(d_A, lda, (int)w.M, (int)w.N, nsub, w.panel_rows, p->wsp.colptr, cnt_sub,
                                             p->wsp.vals, idx);
        wsp_deal_kernel<IdxT><<<(unsigned)((L + 7) / 8), 256>>>(p->wsp.colptr, L, p->wsp.vals, idx);
    }
    SPMV_CUDA(cudaGetLastError());
    SPMV_CUDA(cudaDeviceSynchronize());
    return SPMV_OK;
}

} // namespace

// Fills p->wsp.{colptr, vals, idx} and the host-side description `w` (sizes + colptr) that
// configure_wsp needs.
int pack_wsp_device(spmv_plan *p, const float *d_A, int64_t lda, int index_bits_opt, HostWsp &w)
{
    const int64_t M = p->M, N = p->N;
    w.M = M; w.N = N;
    int rc = count_nnz(d_A, lda, M, N, &w.nnz);
    if (rc) return rc;
    wsp_choose_panels(w, w.nnz, index_bits_opt);
    if (w.index_bits == 16 && w.panel_rows >= 65536) return set_error(SPMV_ERR_ARG, "wsp: 16-bit row ids need fewer than 65536 rows");
    const int64_t L = (int64_t)w.panels * N;
    if ((w.nnz + 3 * L) / 4 >= (int64_t)UINT32_MAX) return set_error(SPMV_ERR_UNSUPPORTED, "wsp: more than 2^32 groups");
    const int64_t prows = w.panels > 1 ? w.panel_rows : M;
    const int nsub = (int)std::max<int64_t>(1, (prows + kWspSub - 1) / kWspSub);
    if ((int64_t)w.panels * nsub > 65535) return set_error(SPMV_ERR_UNSUPPORTED, "wsp: too many rows for the device packer");

    DevTmp tmp;
    uint32_t *cnt_sub = nullptr, *cntg = nullptr;
    rc = tmp.get(&cnt_sub, (size_t)w.panels * nsub * N, false);
    if (!rc) rc = tmp.get(&cntg, (size_t)L + 1, true);
    if (rc) return rc;
    if (M > 0 && N > 0) {
        dim3 grid((unsigned)((N + 255) / 256), (unsigned)(w.panels * nsub));
        wsp_count_kernel<<<grid, 256>>>(d_A, lda, (int)M, (int)N, nsub, prows, cnt_sub);
        wsp_list_kernel<<<dim3((unsigned)((N + 255) / 256), (unsigned)w.panels), 256>>>(cnt_sub, (int)N, nsub, w.panels, cntg);
        SPMV_CUDA(cudaGetLastError());
    }
    rc = plan_alloc(p, reinterpret_cast<void **>(&p->wsp.colptr), ((size_t)L + 1) * sizeof(uint32_t), false);
    if (rc) return rc;
    p->device_bytes += (int64_t)((L + 1) * sizeof(uint32_t));
    unsigned long long total = 0;
    rc = exclusive_scan_u32(cntg, p->wsp.colptr, (size_t)L + 1, &total);
    if (rc) return rc;
    w.groups = (int64_t)total;
    w.colptr.resize((size_t)L + 1);
    SPMV_CUDA(cudaMemcpy(w.colptr.data(), p->wsp.colptr, ((size_t)L + 1) * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    w.max_col_groups = 0;
    for (int64_t i = 0; i < L; i++) w.max_col_groups = std::max<int64_t>(w.max_col_groups, (int64_t)w.colptr[i + 1] - w.colptr[i]);
    // the fill kernels see the panel height through panel_rows: M when there is one panel
    HostWsp geo = w;
    geo.panel_rows = prows;
    if (w.index_bits == 16) rc = wsp_build<uint16_t>(p, d_A, lda, geo, nsub, cnt_sub);
    else rc = wsp_build<uint32_t>(p, d_A, lda, geo, nsub, cnt_sub);
    return rc;
}

// Fills p->panel.{off, rel, vals, idx} and the host-side description `h` (sizes + per-row
// statistics) that configure_panel and the traffic accounting need.
int pack_panel_device(spmv_plan *p, const float *d_A, int64_t lda, bool tiled, int slab_cols_opt, HostPanel &h)
{
    const int64_t M = p->M, N = p->N;
    int64_t nnz = 0;
    int rc = count_nnz(d_A, lda, M, N, &nnz);
    if (rc) return rc;
    const int W = slab_cols_opt > 0 ? slab_cols_opt : choose_slab_cols(M, N, nnz);
    if (W < kMinSlabCols || W > kMaxSlabCols || (W & (W - 1))) return set_error(SPMV_ERR_ARG, "panel: bad slab width %d", W);
    h.M = M; h.N = N; h.tiled = tiled; h.slab_cols = W;
    h.index_bits = W == 256 ? 8 : 16;
    h.slabs = (int)((N + W - 1) / W);
    h.row_blocks = (int)((M + kTileRows - 1) / kTileRows);
    h.nnz = nnz;
    const size_t n_off = (size_t)h.slabs * ((size_t)M + 1);
    const long long segs = (long long)h.slabs * M;
    if ((segs + kPackWarps - 1) / kPackWarps > 0x7fffffffLL) return set_error(SPMV_ERR_UNSUPPORTED, "panel: too many row segments");

    DevTmp tmp;
    uint32_t *off_row = nullptr;                          // per-row offsets: the AWSP table, or TCSR's source
    int *stats = nullptr, *flag = nullptr;
    if (tiled) rc = tmp.get(&off_row, n_off, true);
    else {
        rc = plan_alloc(p, reinterpret_cast<void **>(&p->panel.off), std::max<size_t>(n_off, 1) * sizeof(uint32_t), true);
        off_row = p->panel.off;
    }
    if (!rc) rc = tmp.get(&stats, 3 * (size_t)M, true);
    if (!rc) rc = tmp.get(&flag, 1, true);
    if (rc) return rc;
    if (segs > 0) {
        panel_count_kernel<<<(unsigned)((segs + kPackWarps - 1) / kPackWarps), kPackWarps * 32>>>(
            d_A, lda, (int)M, (int)N, W, h.slabs, off_row, stats, stats + M, stats + 2 * M);
        SPMV_CUDA(cudaGetLastError());
    }
    unsigned long long total = 0;
    rc = exclusive_scan_u32(off_row, off_row, n_off, &total);
    if (rc) return rc;
    if (total >= (unsigned long long)UINT32_MAX) return set_error(SPMV_ERR_UNSUPPORTED, "panel: more than 2^32 groups");
    h.groups = (int64_t)total;

    const size_t entries = (size_t)h.groups * 4;
    const size_t idx_bytes = entries * (h.index_bits == 8 ? 1 : 2);
    rc = plan_alloc(p, reinterpret_cast<void **>(&p->panel.vals), entries * sizeof(float), false);
    if (!rc) rc = plan_alloc(p, &p->panel.idx, idx_bytes, false);
    if (rc) return rc;
    p->device_bytes += (int64_t)(entries * sizeof(float) + idx_bytes);
    if (segs > 0 && h.groups > 0) {
        const int warps = W <= 1024 ? kPackWarps : kPackWarps / 2;
        const size_t per_warp = ((size_t)(W + 4) * 6 + 15) & ~(size_t)15;
        const int smem = (int)(per_warp * warps);
        const unsigned grid = (unsigned)((segs + warps - 1) / warps);
        if (h.index_bits == 8) {
            SPMV_CUDA(cudaFuncSetAttribute(panel_fill_kernel<uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            panel_fill_kernel<uint8_t><<<grid, warps * 32, smem>>>(d_A, lda, (int)M, (int)N, W, h.slabs, warps, off_row,
                                                                   p->panel.vals, reinterpret_cast<uint8_t *>(p->panel.idx));
        } else {
            SPMV_CUDA(cudaFuncSetAttribute(panel_fill_kernel<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            panel_fill_kernel<uint16_t><<<grid, warps * 32, smem>>>(d_A, lda, (int)M, (int)N, W, h.slabs, warps, off_row,
                                                                    p->panel.vals, reinterpret_cast<uint16_t *>(p->panel.idx));
        }
        SPMV_CUDA(cudaGetLastError());
    }
    if (tiled) {
        const size_t n_tile = (size_t)h.slabs * (h.row_blocks + 1), n_rel = (size_t)h.slabs * h.row_blocks * kTileRows;
        rc = plan_alloc(p, reinterpret_cast<void **>(&p->panel.off), std::max<size_t>(n_tile, 1) * sizeof(uint32_t), true);
        if (!rc) rc = plan_alloc(p, reinterpret_cast<void **>(&p->panel.rel), std::max<size_t>(n_rel, 1) * sizeof(uint16_t), true);
        if (rc) return rc;
        if (n_rel) {
            panel_tile_offsets_kernel<<<(unsigned)((n_rel + 255) / 256), 256>>>(off_row, (int)M, h.row_blocks, h.slabs,
                                                                               p->panel.off, p->panel.rel, flag);
            SPMV_CUDA(cudaGetLastError());
        }
        p->device_bytes += (int64_t)(n_tile * 4 + n_rel * 2);
        p->off_bytes = (int64_t)(n_tile * 4 + n_rel * 2);
    } else {
        p->device_bytes += (int64_t)(n_off * 4);
        p->off_bytes = (int64_t)(n_off * 4);
    }
    int flag_h = 0;
    SPMV_CUDA(cudaMemcpy(&flag_h, flag, sizeof flag_h, cudaMemcpyDeviceToHost));
    if (flag_h) return set_error(SPMV_ERR_UNSUPPORTED, "tcsr: a 32-row tile holds more than 65535 groups");
    h.row_nnz.resize((size_t)M); h.row_groups.resize((size_t)M); h.row_segs.resize((size_t)M);
    if (M > 0) {
        SPMV_CUDA(cudaMemcpy(h.row_nnz.data(), stats, (size_t)M * 4, cudaMemcpyDeviceToHost));
        SPMV_CUDA(cudaMemcpy(h.row_groups.data(), stats + M, (size_t)M * 4, cudaMemcpyDeviceToHost));
        SPMV_CUDA(cudaMemcpy(h.row_segs.data(), stats + 2 * M, (size_t)M * 4, cudaMemcpyDeviceToHost));
    }
    h.nonempty_segments = 0;
    for (int32_t v : h.row_segs) h.nonempty_segments += v;
    SPMV_CUDA(cudaDeviceSynchronize());
    return SPMV_OK;
}

// ------------------------------------------------------------------------------------------
// row strips (formats.hpp: HostStrips) from CSR(A^T) ALREADY IN DEVICE MEMORY — the route for
// matrices that never exist densely (BASELINE config 5).  check + count, per-segment counts
// (integer atomics: the counts do not depend on their order), pad to 32 bytes, the shared scan,
// then the fill: one CTA per (band, strip) walks the strip's columns IN ORDER, the threads of
// the CTA share one column's entries — their rows are distinct, so no two threads touch the
// same segment cursor — with a barrier between columns.  Entries of a segment therefore land in
// ascending column order: the bytes are the host packer's (pack_strips_csc).
// ------------------------------------------------------------------------------------------
namespace {

__global__ void __launch_bounds__(256)
strips_check_kernel(const int64_t *__restrict__ col_ptr, const int32_t *__restrict__ row_idx, const float *__restrict__ values,
                    int M, long long N, unsigned long long *__restrict__ nnz, int *__restrict__ bad)
{
    __shared__ unsigned long long ws[8];
    const int lane = threadIdx.x & 31;
    const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), n_warps = (long long)gridDim.x * (blockDim.x >> 5);
    unsigned long long n = 0;
    for (long long c = warp; c < N; c += n_warps) {
        const long long a = col_ptr[c], b = col_ptr[c + 1];
        if (b < a) { if (lane == 0) *bad = 1; continue; }
        for (long long k = a + lane; k < b; k += 32) {
            const int r = row_idx[k];
            if (r < 0 || r >= M || (k > a && row_idx[k - 1] >= r)) *bad = 1;
            n += values[k] != 0.0f;
        }
    }
    n = block_sum_u64(n, ws);
    if (threadIdx.x == 0 && n) atomicAdd(nnz, n);
}

__global__ void __launch_bounds__(256)
strips_count_kernel(const int64_t *__restrict__ col_ptr, const int32_t *__restrict__ row_idx, const float *__restrict__ values,
                    long long M, long long N, int sw, uint32_t *__restrict__ cnt)
{
    const int lane = threadIdx.x & 31;
    const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), n_warps = (long long)gridDim.x * (blockDim.x >> 5);
    const long long band_cols = (long long)sw * kStripsPerBand;
    for (long long c = warp; c < N; c += n_warps) {
        const long long band = c / band_cols;
        const int strip = (int)((c - band * band_cols) / sw);
        uint32_t *base = cnt + (size_t)band * M * kStripsPerBand + strip;
        for (long long k = col_ptr[c] + lane; k < col_ptr[c + 1]; k += 32)
            if (values[k] != 0.0f) atomicAdd(base + (size_t)row_idx[k] * kStripsPerBand, 1u);
    }
}

// counts -> padded counts (in place); per-row statistics for the traffic accounting
__global__ void __launch_bounds__(256)
strips_pad_kernel(uint32_t *__restrict__ cnt, size_t segs, long long M, int *__restrict__ row_nnz, int *__restrict__ row_groups)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= segs) return;
    const uint32_t n = cnt[i], padded = (n + kStripPad - 1) / kStripPad * kStripPad;
    cnt[i] = padded;
    if (n) {
        const long long row = (long long)((i / kStripsPerBand) % (size_t)M);
        atomicAdd(row_nnz + row, (int)n);
        atomicAdd(row_groups + row, (int)(padded / kStripPad));
    }
}

// row-major starts (+ sentinel) -> the strip-major device table [band][strip 0..16][row]
__global__ void __launch_bounds__(256)
strips_transpose_kernel(const uint32_t *__restrict__ row_major, long long M, int bands, uint32_t *__restrict__ strip_major)
{
    const size_t n = (size_t)bands * (kStripsPerBand + 1) * M;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const size_t row = i % (size_t)M, bk = i / (size_t)M;
    const size_t k = bk % (kStripsPerBand + 1), b = bk / (kStripsPerBand + 1);
    strip_major[i] = row_major[(b * (size_t)M + row) * kStripsPerBand + k];
}

__global__ void __launch_bounds__(256)
strips_fill_kernel(const int64_t *__restrict__ col_ptr, const int32_t *__restrict__ row_idx, const float *__restrict__ values,
                   long long M, long long N, int sw, uint32_t *__restrict__ cursor, uint2 *__restrict__ ent)
{
    const long long band = blockIdx.x / kStripsPerBand;
    const int strip = blockIdx.x % kStripsPerBand;
    const long long c0 = (band * kStripsPerBand + strip) * (long long)sw, c1 = min(N, c0 + sw);
    uint32_t *cur = cursor + (size_t)band * M * kStripsPerBand + strip;
    for (long long c = c0; c < c1; c++) {
        const long long a = col_ptr[c], b = col_ptr[c + 1];
        for (long long k = a + threadIdx.x; k < b; k += blockDim.x) {
            const float v = values[k];
            if (v != 0.0f) {
                uint32_t *q = cur + (size_t)row_idx[k] * kStripsPerBand;   // rows of one column are distinct: no other thread is here
                const uint32_t p = *q;
                *q = p + 1u;
                ent[p] = make_uint2(__float_as_uint(v), (uint32_t)(c - c0) + 1u);
            }
        }
        __syncthreads();                                  // the next column may touch the same cursors
    }
}

} // namespace

int pack_strips_csc_device(spmv_plan *p, const int64_t *d_col_ptr, const int32_t *d_row_idx, const float *d_values,
                           int strip_cols_opt, HostStrips &h)
{
    const int64_t M = p->M, N = p->N;
    DevTmp tmp;
    unsigned long long *nnz_d = nullptr;
    int *bad = nullptr, *stats = nullptr;
    int rc = tmp.get(&nnz_d, 1, true);
    if (!rc) rc = tmp.get(&bad, 1, true);
    if (!rc) rc = tmp.get(&stats, 2 * (size_t)std::max<int64_t>(M, 1), true);
    if (rc) return rc;
    const unsigned col_blocks = (unsigned)std::min<int64_t>(std::max<int64_t>((N + 7) / 8, 1), 148 * 16);
    strips_check_kernel<<<col_blocks, 256>>>(d_col_ptr, d_row_idx, d_values, (int)M, (long long)N, nnz_d, bad);
    SPMV_CUDA(cudaGetLastError());
    unsigned long long nnz = 0;
    int bad_h = 0;
    SPMV_CUDA(cudaMemcpy(&nnz, nnz_d, sizeof nnz, cudaMemcpyDeviceToHost));
    SPMV_CUDA(cudaMemcpy(&bad_h, bad, sizeof bad_h, cudaMemcpyDeviceToHost));
    if (bad_h) return set_error(SPMV_ERR_ARG, "CSR(A^T) input: col_ptr must be monotone, the rows of a column in range and strictly ascending");
    if (nnz >= (unsigned long long)UINT32_MAX) return set_error(SPMV_ERR_UNSUPPORTED, "strips: more than 2^32 entries");
    const int sw = strip_cols_opt > 0 ? strip_cols_opt : choose_strip_cols(M, N, (int64_t)nnz);
    if (sw < 32 || sw > kMaxStripCols || sw % 32) return set_error(SPMV_ERR_ARG, "strips: bad strip width %d", sw);
    h.M = M; h.N = N; h.nnz = (int64_t)nnz; h.strip_cols = sw;
    const int64_t band_cols = (int64_t)sw * kStripsPerBand;
    h.bands = (int)std::max<int64_t>(1, (N + band_cols - 1) / band_cols);
    const size_t segs = (size_t)h.bands * M * kStripsPerBand;
    if (segs + 1 >= (size_t)UINT32_MAX * 8) return set_error(SPMV_ERR_UNSUPPORTED, "strips: too many row segments");

    uint32_t *soff = nullptr;                             // row-major counts -> starts (+ sentinel): a temporary
    rc = tmp.get(&soff, segs + 1, true);
    if (rc) return rc;
    if (segs && nnz) {
        strips_count_kernel<<<col_blocks, 256>>>(d_col_ptr, d_row_idx, d_values, (long long)M, (long long)N, sw, soff);
        strips_pad_kernel<<<(unsigned)((segs + 255) / 256), 256>>>(soff, segs, (long long)M, stats, stats + M);
        SPMV_CUDA(cudaGetLastError());
    }
    unsigned long long total = 0;
    rc = exclusive_scan_u32(soff, soff, segs + 1, &total);
    if (rc) return rc;
    if (total >= (unsigned long long)UINT32_MAX) return set_error(SPMV_ERR_UNSUPPORTED, "strips: more than 2^32 stored entries");
    // + 32 spare entries: an idle lane's (unread) source address stays legal (capi.cu: setup_strips)
    rc = plan_alloc(p, reinterpret_cast<void **>(&p->strips.ent), ((size_t)total + 32) * sizeof(uint2), true);
    if (rc) return rc;
    if (total) {
        uint32_t *cursor = nullptr;
        rc = tmp.get(&cursor, segs, false);
        if (rc) return rc;
        SPMV_CUDA(cudaMemcpy(cursor, soff, segs * sizeof(uint32_t), cudaMemcpyDeviceToDevice));
        strips_fill_kernel<<<(unsigned)(h.bands * kStripsPerBand), 256>>>(d_col_ptr, d_row_idx, d_values, (long long)M, (long long)N, sw,
                                                                         cursor, p->strips.ent);
        SPMV_CUDA(cudaGetLastError());
    }
    const size_t n_dev_off = (size_t)h.bands * (kStripsPerBand + 1) * M;
    rc = plan_alloc(p, reinterpret_cast<void **>(&p->strips.soff), (n_dev_off + 1) * sizeof(uint32_t), true);
    if (rc) return rc;
    if (n_dev_off) {
        strips_transpose_kernel<<<(unsigned)((n_dev_off + 255) / 256), 256>>>(soff, (long long)M, h.bands, p->strips.soff);
        SPMV_CUDA(cudaGetLastError());
    }
    p->device_bytes += (int64_t)((n_dev_off + 1) * 4 + ((size_t)total + 32) * 8);
    p->off_bytes = (int64_t)((n_dev_off + 1) * 4);
    p->fmt_groups = (int64_t)total / kStripPad;
    h.row_nnz.assign((size_t)M, 0); h.row_groups.assign((size_t)M, 0);
    if (M > 0) {
        SPMV_CUDA(cudaMemcpy(h.row_nnz.data(), stats, (size_t)M * 4, cudaMemcpyDeviceToHost));
        SPMV_CUDA(cudaMemcpy(h.row_groups.data(), stats + M, (size_t)M * 4, cudaMemcpyDeviceToHost));
    }
    SPMV_CUDA(cudaDeviceSynchronize());
    return SPMV_OK;
}


} // namespace spmv
