// pack_host.cpp — host packers.
//   (1) this library's device formats (formats.hpp) from a dense matrix or from CSR(A^T);
//   (2) the reference's six host layouts, bit-exact, behind spmv_ref_pack() — what the drop-in
//       format classes (host/formats.cpp) are made of.
// Written from the layout descriptions in SURVEY.md §2a; checked bit-for-bit against the
// reference packers through the oracle (tests/test_oracle_pinned.py, tests/test_packers_cpu.py).
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <new>
#include <thread>

#include "formats.hpp"
#include "deal.hpp"
#include "plan.hpp"
#include "spmv_b200.h"

namespace spmv {

// CSR(A^T) input rule (spmv_b200.h): inside a column the rows ascend STRICTLY — a repeated (row,
// column) pair would make two entries of one segment share an accumulator in the panel and strip
// kernels (lost update) while wsp would add them, so the variants would disagree; out-of-range rows
// and unsorted lists are rejected the same way.
int check_csc(int64_t M, int64_t N, const int64_t *col_ptr, const int32_t *row_idx)
{
    for (int64_t i = 0; i < N; i++) {
        if (col_ptr[i + 1] < col_ptr[i]) return SPMV_ERR_ARG;
        for (int64_t k = col_ptr[i]; k < col_ptr[i + 1]; k++) {
            if (row_idx[k] < 0 || row_idx[k] >= M) return SPMV_ERR_ARG;
            if (k > col_ptr[i] && row_idx[k] <= row_idx[k - 1]) return SPMV_ERR_ARG;
        }
    }
    return SPMV_OK;
}

// ---------------------------------------------------------------------------- WSP ---------
// Lists are indexed by l = panel * N + column.
static void wsp_finish_layout(HostWsp &w, const std::vector<int64_t> &list_nnz)
{
    const int64_t L = (int64_t)w.panels * w.N;
    w.colptr.resize(L + 1);
    int64_t g = 0, mx = 0;
    for (int64_t i = 0; i < L; i++) {
        w.colptr[i] = (uint32_t)g;
        int64_t cg = (list_nnz[i] + 3) / 4;
        mx = std::max(mx, cg);
        g += cg;
    }
    w.colptr[L] = (uint32_t)g;
    w.groups = g;
    w.max_col_groups = mx;
    const uint32_t pad = (uint32_t)w.panel_rows;
    // one spare pad group at the end: the ring kernel may address group `groups` (never uses it)
    w.vals.assign((size_t)(g + 1) * 4, 0.0f);
    if (w.index_bits == 16) w.idx16.assign((size_t)(g + 1) * 4, (uint16_t)pad);
    else w.idx32.assign((size_t)(g + 1) * 4, pad);
}

static int pack_threads(int64_t work_items)
{
    int n = (int)std::min<int64_t>(std::max(1u, std::thread::hardware_concurrency()), 16);
    if (const char *e = std::getenv("SPMV_PACK_THREADS")) n = std::max(1, std::atoi(e));
    return (int)std::max<int64_t>(1, std::min<int64_t>(n, work_items));
}

// In-chunk order for the kernel's shared-memory gathers of x (deal.hpp).  Lists are independent:
// a small pool of host threads takes contiguous ranges of them (same bytes for any thread count).
template <class IdxT> static void wsp_bank_deal(HostWsp &w, std::vector<IdxT> &idx)
{
    const int64_t L = (int64_t)w.panels * w.N;
    auto deal_lists = [&](int64_t l0, int64_t l1) {
        IdxT oi[128], li[128];
        float ov[128], lv[128];
        for (int64_t c = l0; c < l1; c++)
            for (int64_t c0 = w.colptr[c]; c0 < w.colptr[c + 1]; c0 += 32) {
                const int lanes = (int)std::min<int64_t>(32, w.colptr[c + 1] - c0);
                const size_t first = (size_t)c0 * 4;
                const int count = 4 * lanes;
                std::copy(idx.begin() + first, idx.begin() + first + count, li);
                std::copy(w.vals.begin() + first, w.vals.begin() + first + count, lv);
                deal_chunk(lanes, [&](int k) { return (unsigned)li[k]; },
                           [&](int slot, int k) { oi[slot] = li[k]; ov[slot] = lv[k]; });
                std::copy(oi, oi + count, idx.begin() + first);
                std::copy(ov, ov + count, w.vals.begin() + first);
            }
    };
    const int n_threads = w.groups < 4096 ? 1 : pack_threads(L);
    if (n_threads == 1) { deal_lists(0, L); return; }
    // ranges of about equal group counts (lists may be very uneven: config 4)
    std::vector<int64_t> cut((size_t)n_threads + 1, L);
    cut[0] = 0;
    int64_t c = 0;
    for (int t = 1; t < n_threads; t++) {
        const int64_t target = w.groups * t / n_threads;
        while (c < L && (int64_t)w.colptr[c] < target) c++;
        cut[(size_t)t] = c;
    }
    std::vector<std::thread> pool;
    for (int t = 0; t < n_threads; t++) pool.emplace_back(deal_lists, cut[(size_t)t], cut[(size_t)t + 1]);
    for (std::thread &th : pool) th.join();
}

static void wsp_bank_order(HostWsp &w)
{
    if (w.index_bits == 16) wsp_bank_deal(w, w.idx16); else wsp_bank_deal(w, w.idx32);
}

// Row panels: when x (M floats) does not fit shared memory next to the ring but the lists
// stay long enough after the cut (>= 32 non-zeros per (panel, column) on average), cut the rows
// into panels of 12288 (48 KB of x + 48 KB of ring per CTA: two CTAs per SM).
void wsp_choose_panels(HostWsp &w, int64_t nnz, int index_bits_opt)
{
    constexpr int64_t kPanelRows = 12288;
    w.panels = 1; w.panel_rows = w.M;
    const bool x_fits = ((size_t)w.M + 4) * sizeof(float) <= 96 * 1024;
    if (!x_fits && index_bits_opt != 32 && w.N > 0) {
        const int64_t P = (w.M + kPanelRows - 1) / kPanelRows;
        if ((double)nnz / ((double)P * (double)w.N) >= 32.0) { w.panels = (int)P; w.panel_rows = kPanelRows; }
    }
    w.index_bits = index_bits_opt ? index_bits_opt : (w.panel_rows < 65536 ? 16 : 32);
}

int pack_wsp_dense(int64_t M, int64_t N, const float *A, int64_t lda, int index_bits, HostWsp &w)
{
    w.M = M; w.N = N;
    std::vector<int64_t> col_cnt((size_t)N, 0);
    int64_t nnz = 0;
    for (int64_t j = 0; j < M; j++) {
        const float *row = A + j * lda;
        for (int64_t i = 0; i < N; i++) nnz += (row[i] != 0.0f);
    }
    w.nnz = nnz;
    wsp_choose_panels(w, nnz, index_bits);
    if (w.index_bits == 16 && w.panel_rows >= 65536) return SPMV_ERR_ARG;
    std::vector<int64_t> cnt((size_t)w.panels * N, 0);
    for (int64_t j = 0; j < M; j++) {            // row-major sweep: rows arrive in ascending order
        const float *row = A + j * lda;
        int64_t *c = cnt.data() + (w.panels > 1 ? j / w.panel_rows : 0) * N;
        for (int64_t i = 0; i < N; i++) c[i] += (row[i] != 0.0f);
    }
    if ((nnz + 3 * (int64_t)cnt.size()) / 4 >= (int64_t)UINT32_MAX) return SPMV_ERR_UNSUPPORTED;
    wsp_finish_layout(w, cnt);
    std::vector<int64_t> cur(cnt.size());
    for (size_t i = 0; i < cnt.size(); i++) cur[i] = (int64_t)w.colptr[i] * 4;
    for (int64_t j = 0; j < M; j++) {
        const float *row = A + j * lda;
        const int64_t pj = w.panels > 1 ? j / w.panel_rows : 0;
        const int64_t local = j - pj * (w.panels > 1 ? w.panel_rows : 0);
        int64_t *c = cur.data() + pj * N;
        for (int64_t i = 0; i < N; i++) {
            float v = row[i];
            if (v != 0.0f) {
                int64_t p = c[i]++;
                w.vals[p] = v;
                if (w.index_bits == 16) w.idx16[p] = (uint16_t)local; else w.idx32[p] = (uint32_t)local;
            }
        }
    }
    wsp_bank_order(w);
    return SPMV_OK;
}

int pack_wsp_csc(int64_t M, int64_t N, const int64_t *col_ptr, const int32_t *row_idx,
                 const float *values, int index_bits, HostWsp &w)
{
    w.M = M; w.N = N;
    int64_t nnz = 0;
    for (int64_t k = col_ptr[0]; k < col_ptr[N]; k++) {
        if (row_idx[k] < 0 || row_idx[k] >= M) return SPMV_ERR_ARG;
        nnz += (values[k] != 0.0f);
    }
    w.nnz = nnz;
    wsp_choose_panels(w, nnz, index_bits);
    if (w.index_bits == 16 && w.panel_rows >= 65536) return SPMV_ERR_ARG;
    std::vector<int64_t> cnt((size_t)w.panels * N, 0);
    for (int64_t i = 0; i < N; i++)
        for (int64_t k = col_ptr[i]; k < col_ptr[i + 1]; k++)
            if (values[k] != 0.0f) cnt[(size_t)(w.panels > 1 ? row_idx[k] / w.panel_rows : 0) * N + i]++;
    if ((nnz + 3 * (int64_t)cnt.size()) / 4 >= (int64_t)UINT32_MAX) return SPMV_ERR_UNSUPPORTED;
    wsp_finish_layout(w, cnt);
    std::vector<int64_t> cur(cnt.size());
    for (size_t i = 0; i < cnt.size(); i++) cur[i] = (int64_t)w.colptr[i] * 4;
    for (int64_t i = 0; i < N; i++)
        for (int64_t k = col_ptr[i]; k < col_ptr[i + 1]; k++) {
            if (values[k] == 0.0f) continue;
            const int64_t pj = w.panels > 1 ? row_idx[k] / w.panel_rows : 0;
            const int64_t local = row_idx[k] - pj * (w.panels > 1 ? w.panel_rows : 0);
            const int64_t p = cur[(size_t)pj * N + i]++;
            w.vals[p] = values[k];
            if (w.index_bits == 16) w.idx16[p] = (uint16_t)local; else w.idx32[p] = (uint32_t)local;
        }
    wsp_bank_order(w);
    return SPMV_OK;
}

// ---------------------------------------------------------------------------- panel -------
namespace {

// Row source abstraction: yields the non-zeros of (slab, row) in ascending column order.
struct DenseSource {
    const float *A; int64_t lda, N; int W;
    void begin_slab(int) {}
    int get(int slab, int64_t row, uint16_t *cols, float *vals) const
    {
        const int64_t c0 = (int64_t)slab * W;
        const int cw = (int)std::min<int64_t>(W, N - c0);
        const float *p = A + row * lda + c0;
        int n = 0;
        for (int c = 0; c < cw; c++) {
            const float v = p[c];
            if (v != 0.0f) { cols[n] = (uint16_t)c; vals[n] = v; n++; }
        }
        return n;
    }
};

struct CscSource {
    const int64_t *col_ptr; const int32_t *row_idx; const float *values; int64_t M, N; int W;
    // per-slab row buckets
    std::vector<int64_t> start;    // M+1
    std::vector<uint16_t> col;     // local column of each entry, row-bucketed
    std::vector<float> val;
    void begin_slab(int slab)
    {
        const int64_t c0 = (int64_t)slab * W;
        const int cw = (int)std::min<int64_t>(W, N - c0);
        start.assign((size_t)M + 1, 0);
        for (int c = 0; c < cw; c++)
            for (int64_t k = col_ptr[c0 + c]; k < col_ptr[c0 + c + 1]; k++)
                if (values[k] != 0.0f) start[(size_t)row_idx[k] + 1]++;
        for (int64_t r = 0; r < M; r++) start[r + 1] += start[r];
        col.resize((size_t)start[M]); val.resize((size_t)start[M]);
        std::vector<int64_t> cur(start.begin(), start.end() - 1);
        for (int c = 0; c < cw; c++)      // ascending columns => column order inside each row
            for (int64_t k = col_ptr[c0 + c]; k < col_ptr[c0 + c + 1]; k++)
                if (values[k] != 0.0f) {
                    const int64_t p = cur[row_idx[k]]++;
                    col[p] = (uint16_t)c; val[p] = values[k];
                }
    }
    int get(int, int64_t row, uint16_t *cols, float *vals) const
    {
        const int64_t b = start[row], e = start[row + 1];
        for (int64_t p = b; p < e; p++) { cols[p - b] = col[p]; vals[p - b] = val[p]; }
        return (int)(e - b);
    }
};

// One slab's share of the panel format, built independently of the other slabs (offsets are
// relative to the slab's first group); pack_panel concatenates the slabs.
struct SlabOut {
    std::vector<float> vals;
    std::vector<uint8_t> idx8;
    std::vector<uint16_t> idx16;
    std::vector<uint32_t> off;      // per_slab entries, slab-relative
    std::vector<uint16_t> rel;      // tiled only
    int64_t groups = 0, nnz = 0, segs = 0;
    int rc = SPMV_OK;
};

struct RowStats { std::vector<int32_t> nnz, groups, segs; };

template <class Source>
void pack_slab(int s, int64_t M, bool tiled, int W, int index_bits, int row_blocks, Source &src, SlabOut &O, RowStats &st)
{
    const int64_t per_slab = tiled ? (row_blocks + 1) : (M + 1);
    O.off.assign((size_t)per_slab, 0);
    if (tiled) O.rel.assign((size_t)row_blocks * kTileRows, 0);
    std::vector<uint16_t> cols((size_t)W + 4), ocols((size_t)W);
    std::vector<float> vals((size_t)W + 4), ovals((size_t)W);

    // In-chunk order for the kernel's shared-memory accumulators (deal.hpp).  Pads carry value 0
    // and the smallest column ABSENT from the segment: they add an exact 0 to an accumulator no
    // real entry of this row touches (n % 4 != 0 implies n < W).
    auto emit = [&](int n) {
        const int g = (n + 3) / 4;
        const size_t at = O.vals.size();
        uint16_t absent = 0;
        for (int k = 0; k < n && cols[k] == absent; k++) absent++;
        for (int k = n; k < 4 * g; k++) { cols[k] = absent; vals[k] = 0.0f; }
        for (int c0 = 0; c0 < g; c0 += 32) {               // one chunk = up to 32 groups
            const int lanes = std::min(32, g - c0);
            const int first = 4 * c0;
            deal_chunk(lanes, [&](int k) { return (unsigned)cols[first + k]; },
                       [&](int slot, int k) { ocols[first + slot] = cols[first + k]; ovals[first + slot] = vals[first + k]; });
        }
        O.vals.resize(at + (size_t)g * 4);
        std::memcpy(&O.vals[at], ovals.data(), sizeof(float) * (size_t)g * 4);
        if (index_bits == 8) {
            O.idx8.resize(at + (size_t)g * 4);
            for (int k = 0; k < 4 * g; k++) O.idx8[at + k] = (uint8_t)ocols[k];
        } else {
            O.idx16.resize(at + (size_t)g * 4);
            std::memcpy(&O.idx16[at], ocols.data(), sizeof(uint16_t) * (size_t)g * 4);
        }
        O.groups += g;
        return g;
    };

    src.begin_slab(s);
    for (int rb = 0; rb < row_blocks; rb++) {
        const int64_t tile_first = O.groups;
        if (tiled) O.off[(size_t)rb] = (uint32_t)tile_first;
        for (int r = 0; r < kTileRows; r++) {
            const int64_t row = (int64_t)rb * kTileRows + r;
            if (tiled) O.rel[(size_t)rb * kTileRows + r] = (uint16_t)(O.groups - tile_first);
            if (row >= M) continue;
            if (!tiled) O.off[(size_t)row] = (uint32_t)O.groups;
            const int n = src.get(s, row, cols.data(), vals.data());
            if (n) {
                const int g = emit(n);
                st.nnz[row] += n; st.groups[row] += g; st.segs[row] += 1; O.nnz += n; O.segs++;
            }
        }
        if (O.groups - tile_first > 65535) { O.rc = SPMV_ERR_UNSUPPORTED; return; }   // u16 rel offsets
        if (O.groups >= (int64_t)UINT32_MAX) { O.rc = SPMV_ERR_UNSUPPORTED; return; }
    }
    O.off[(size_t)(tiled ? row_blocks : M)] = (uint32_t)O.groups;
}

// Lane-owned blocks (formats.hpp): per (slab, block) one stream per lane, padded to the longest.
template <class Source>
void pack_slab_lob(int s, int64_t M, int W, int R, Source &src, SlabOut &O, RowStats &st)
{
    int cbits = 0;
    while ((32 << cbits) < W) cbits++;
    const int64_t NB = (M + R - 1) / R;
    O.off.assign((size_t)NB + 1, 0);
    std::vector<uint16_t> cols((size_t)W + 4), li[32];
    std::vector<float> vals((size_t)W + 4), lv[32];
    std::vector<int64_t> lp[32];                           // position of every entry inside its lane's stream
    src.begin_slab(s);
    for (int64_t b = 0; b < NB; b++) {
        O.off[(size_t)b] = (uint32_t)O.groups;
        for (int l = 0; l < 32; l++) { li[l].clear(); lv[l].clear(); }
        const int64_t r1 = std::min<int64_t>(M, (b + 1) * R);
        for (int64_t row = b * R; row < r1; row++) {
            const int n = src.get(s, row, cols.data(), vals.data());
            for (int k = 0; k < n; k++) {
                const int l = cols[k] & 31;
                li[l].push_back((uint16_t)(((row - b * R) << cbits) | (cols[k] >> 5)));
                lv[l].push_back(vals[k]);
            }
            if (n) { st.nnz[row] += n; st.groups[row] += (n + 3) / 4; st.segs[row] += 1; O.nnz += n; O.segs++; }
        }
        size_t longest = 0;
        for (int l = 0; l < 32; l++) longest = std::max(longest, li[l].size());
        // Row-aligned streams: an entry of row r is never placed before chunk floor(r * G0 / span) - slack,
        // so a lane with few entries gets its pads in between rather than at the end and all 32
        // lanes of a chunk look at rows a few apart — their x look-ups (x slice in shared memory,
        // bank = row mod 32) then fall into distinct banks.
        const int64_t G0 = (int64_t)((longest + 3) / 4);
        const int64_t span = r1 - b * R;
        constexpr int64_t slack = 6;   // chunks an entry may run ahead of its row: 8.9 % padding and 1.8 wavefronts per
                                       // look-up on config 5 (0: 12.2 % / 1.4; unaligned streams: 8.5 % / 2.4)
        int64_t G = G0;
        for (int l = 0; l < 32; l++) {
            lp[l].resize(li[l].size());
            int64_t prev = -1;
            for (size_t k = 0; k < li[l].size(); k++) {
                const int64_t row = li[l][k] >> cbits;
                prev = std::max(prev + 1, 4 * std::max<int64_t>(0, row * G0 / span - slack));
                lp[l][k] = prev;
            }
            if (prev >= 0) G = std::max(G, prev / 4 + 1);
        }
        const size_t at = O.vals.size();
        O.vals.resize(at + (size_t)G * 128);
        O.idx16.resize(at + (size_t)G * 128);
        for (int l = 0; l < 32; l++) {
            // a pad repeats the next real entry of the lane (the last one after the end) with
            // value 0: same accumulator, a row that exists and lies in the chunk's row window
            size_t k = 0;
            for (int64_t pos = 0; pos < 4 * G; pos++) {
                const size_t o = at + ((size_t)(pos >> 2) * 32 + l) * 4 + (pos & 3);
                if (k < li[l].size() && lp[l][k] == pos) { O.vals[o] = lv[l][k]; O.idx16[o] = li[l][k]; k++; }
                else {
                    O.vals[o] = 0.0f;
                    O.idx16[o] = li[l].empty() ? (uint16_t)0 : li[l][std::min(k, li[l].size() - 1)];
                }
            }
        }
        O.groups += G * 32;
        if (O.groups >= (int64_t)UINT32_MAX) { O.rc = SPMV_ERR_UNSUPPORTED; return; }
    }
    O.off[(size_t)NB] = (uint32_t)O.groups;
}

// Slabs are independent, so they are packed by a small pool of host threads (the reference packs
// on one thread with vector<vector<float>> copies, awsp.cpp:30-46) and concatenated afterwards;
// the result does not depend on the thread count.
template <class Source>
int pack_panel(int64_t M, int64_t N, bool tiled, int W, bool lane_owned, const Source &proto, HostPanel &P)
{
    if (W < kMinSlabCols || W > kMaxSlabCols || (W & (W - 1))) return SPMV_ERR_ARG;
    if (lane_owned) tiled = false;
    if (lane_owned && W < 1024) return SPMV_ERR_ARG;       // blocks of at most 2048 rows (the kernel stages their x slices)
    P.M = M; P.N = N; P.tiled = tiled;
    P.slab_cols = W;
    P.index_bits = (W == 256 && !lane_owned) ? 8 : 16;
    P.block_rows = 0; P.lob_blocks = 0;
    if (lane_owned) {
        int cbits = 0;
        while ((32 << cbits) < W) cbits++;
        P.block_rows = std::min(kLobMaxBlockRows, 1 << (16 - cbits));
        P.lob_blocks = (int)((M + P.block_rows - 1) / P.block_rows);
    }
    P.slabs = (int)((N + W - 1) / W);
    P.row_blocks = (int)((M + kTileRows - 1) / kTileRows);
    P.nnz = 0; P.groups = 0; P.nonempty_segments = 0;
    P.vals.clear(); P.idx8.clear(); P.idx16.clear(); P.rel.clear();
    const int64_t per_slab = lane_owned ? (P.lob_blocks + 1) : tiled ? (P.row_blocks + 1) : (M + 1);

    const int n_threads = pack_threads(P.slabs);
    std::vector<SlabOut> outs((size_t)P.slabs);
    std::vector<RowStats> stats((size_t)n_threads);
    for (RowStats &st : stats) { st.nnz.assign((size_t)M, 0); st.groups.assign((size_t)M, 0); st.segs.assign((size_t)M, 0); }
    auto worker = [&](int t) {
        Source src = proto;                                // per-thread scratch (CSR(A^T) row buckets)
        for (int s = t; s < P.slabs; s += n_threads) {
            if (lane_owned) pack_slab_lob(s, M, W, P.block_rows, src, outs[(size_t)s], stats[(size_t)t]);
            else pack_slab(s, M, tiled, W, P.index_bits, P.row_blocks, src, outs[(size_t)s], stats[(size_t)t]);
        }
    };
    if (n_threads == 1) worker(0);
    else {
        std::vector<std::thread> pool;
        for (int t = 0; t < n_threads; t++) pool.emplace_back(worker, t);
        for (std::thread &th : pool) th.join();
    }

    int64_t total = 0;
    for (const SlabOut &o : outs) { if (o.rc) return o.rc; total += o.groups; }
    if (total >= (int64_t)UINT32_MAX) return SPMV_ERR_UNSUPPORTED;
    P.vals.resize((size_t)total * 4);
    if (P.index_bits == 8) P.idx8.resize((size_t)total * 4); else P.idx16.resize((size_t)total * 4);
    P.off.assign((size_t)P.slabs * per_slab, 0);
    if (tiled) P.rel.assign((size_t)P.slabs * P.row_blocks * kTileRows, 0);
    int64_t base = 0;
    for (int s = 0; s < P.slabs; s++) {
        const SlabOut &o = outs[(size_t)s];
        if (o.groups) {
            std::memcpy(&P.vals[(size_t)base * 4], o.vals.data(), sizeof(float) * (size_t)o.groups * 4);
            if (P.index_bits == 8) std::memcpy(&P.idx8[(size_t)base * 4], o.idx8.data(), (size_t)o.groups * 4);
            else std::memcpy(&P.idx16[(size_t)base * 4], o.idx16.data(), sizeof(uint16_t) * (size_t)o.groups * 4);
        }
        for (int64_t k = 0; k < per_slab; k++) P.off[(size_t)s * per_slab + k] = (uint32_t)(base + o.off[(size_t)k]);
        if (tiled) std::memcpy(&P.rel[(size_t)s * P.row_blocks * kTileRows], o.rel.data(), sizeof(uint16_t) * o.rel.size());
        base += o.groups;
        P.nnz += o.nnz; P.nonempty_segments += o.segs;
        outs[(size_t)s] = SlabOut();                       // release the slab's copy early
    }
    P.groups = total;
    P.row_nnz.assign((size_t)M, 0); P.row_groups.assign((size_t)M, 0); P.row_segs.assign((size_t)M, 0);
    for (const RowStats &st : stats)
        for (int64_t r = 0; r < M; r++) { P.row_nnz[r] += st.nnz[r]; P.row_groups[r] += st.groups[r]; P.row_segs[r] += st.segs[r]; }
    return SPMV_OK;
}
} // namespace

// Slab width from the density.  DRAM wants long contiguous pieces (measured on B200,
// tools/ubench/bulk_vs_ldgsts.cu: 400-byte segments with equal gaps stream at 3.5 TB/s, >= 1 KB
// pieces at 6.4 TB/s) and full 32-group chunks keep the lanes busy, so aim at ~300 non-zeros per
// row segment; but keep at least ~1000 (slab, 32-row block) work units so a small matrix still
// spreads over the SMs.  256-wide slabs use 8-bit column ids, wider ones 16-bit.
int choose_slab_cols(int64_t M, int64_t N, int64_t nnz)
{
    if (M <= 0 || N <= 0 || nnz <= 0) return kMinSlabCols;
    const double density = (double)nnz / ((double)M * (double)N);
    int w = kMinSlabCols;
    while (w < kMaxSlabCols && w * density < 300.0) w <<= 1;
    // segments that stay short even at the widest slab (under 12 groups: the kernel's multi-row
    // mode, bound by latency rather than bytes): 2048 columns, so that two CTAs fit per SM
    if (w == kMaxSlabCols && w * density < 48.0) w = 2048;
    const int64_t row_blocks = (M + kTileRows - 1) / kTileRows;
    while (w > kMinSlabCols && ((N + w - 1) / w) * row_blocks < 1024) w >>= 1;
    // and at least ~10 slabs: the kernel's CTAs are spread over the slabs, and the last CTA of
    // a slab adds all of that slab's partial rows, so few wide slabs mean a long serial tail
    while (w > kMinSlabCols && N / w < 10) w >>= 1;
    return w;
}

int pack_panel_dense(int64_t M, int64_t N, const float *A, int64_t lda, bool tiled, int slab_cols,
                     HostPanel &P, bool lane_owned)
{
    if (lane_owned && slab_cols <= 0) slab_cols = kLobSlabCols;
    if (slab_cols <= 0) {
        int64_t nnz = 0;
        for (int64_t j = 0; j < M; j++) {
            const float *row = A + j * lda;
            for (int64_t i = 0; i < N; i++) nnz += (row[i] != 0.0f);
        }
        slab_cols = choose_slab_cols(M, N, nnz);
    }
    DenseSource src{A, lda, N, slab_cols};
    return pack_panel(M, N, tiled, slab_cols, lane_owned, src, P);
}

int pack_panel_csc(int64_t M, int64_t N, const int64_t *col_ptr, const int32_t *row_idx,
                   const float *values, bool tiled, int slab_cols, HostPanel &P, bool lane_owned)
{
    if (lane_owned && slab_cols <= 0) slab_cols = kLobSlabCols;
    int64_t nnz = 0;
    for (int64_t k = col_ptr[0]; k < col_ptr[N]; k++) {
        if (row_idx[k] < 0 || row_idx[k] >= M) return SPMV_ERR_ARG;
        nnz += (values[k] != 0.0f);
    }
    if (slab_cols <= 0) slab_cols = choose_slab_cols(M, N, nnz);
    CscSource src{col_ptr, row_idx, values, M, N, slab_cols, {}, {}, {}};
    return pack_panel(M, N, tiled, slab_cols, lane_owned, src, P);
}

// ---------------------------------------------------------------------------- row strips ---
// Strip width from the density: about kStripTargetNnz non-zeros per (row, strip), a whole number
// of bands over N, a multiple of 32 columns, at most kMaxStripCols (shared-memory accumulators).
int choose_strip_cols(int64_t M, int64_t N, int64_t nnz)
{
    if (M <= 0 || N <= 0 || nnz <= 0) return 32;
    const double density = (double)nnz / ((double)M * (double)N);
    const double want = std::min<double>(kMaxStripCols, std::max(32.0, kStripTargetNnz / density));
    const int64_t bands = std::max<int64_t>(1, (int64_t)((double)N / (kStripsPerBand * want) + 0.5));
    int64_t w = (N + bands * kStripsPerBand - 1) / (bands * kStripsPerBand);
    w = (w + 31) / 32 * 32;
    return (int)std::min<int64_t>(kMaxStripCols, std::max<int64_t>(32, w));
}

namespace {
void strips_begin(int64_t M, int64_t N, int strip_cols, HostStrips &h)
{
    h.M = M; h.N = N; h.strip_cols = strip_cols;
    const int64_t band_cols = (int64_t)strip_cols * kStripsPerBand;
    h.bands = (int)std::max<int64_t>(1, (N + band_cols - 1) / band_cols);
    h.soff.assign((size_t)h.bands * M * kStripsPerBand + 1, 0u);
    h.row_nnz.assign((size_t)M, 0);
    h.row_groups.assign((size_t)M, 0);
    h.nnz = 0;
}
inline uint64_t strip_entry(float v, uint32_t col)       // column inside the strip, stored + 1 (formats.hpp)
{
    uint32_t bits;
    std::memcpy(&bits, &v, 4);
    return (uint64_t)bits | ((uint64_t)(col + 1u) << 32);
}
} // namespace

int pack_strips_dense(int64_t M, int64_t N, const float *A, int64_t lda, int strip_cols, HostStrips &h)
{
    if (strip_cols <= 0) {
        int64_t nnz = 0;
        for (int64_t j = 0; j < M; j++) {
            const float *row = A + j * lda;
            for (int64_t i = 0; i < N; i++) nnz += (row[i] != 0.0f);
        }
        strip_cols = choose_strip_cols(M, N, nnz);
    }
    if (strip_cols < 32 || strip_cols > kMaxStripCols || strip_cols % 32) return SPMV_ERR_ARG;
    strips_begin(M, N, strip_cols, h);
    const int64_t band_cols = (int64_t)strip_cols * kStripsPerBand;
    h.ent.clear();
    size_t seg = 0;
    for (int b = 0; b < h.bands; b++)
        for (int64_t j = 0; j < M; j++) {
            const float *row = A + j * lda;
            for (int s = 0; s < kStripsPerBand; s++, seg++) {
                if (h.ent.size() >= (size_t)UINT32_MAX) return SPMV_ERR_UNSUPPORTED;
                h.soff[seg] = (uint32_t)h.ent.size();
                const int64_t c0 = b * band_cols + (int64_t)s * strip_cols;
                const int64_t c1 = std::min<int64_t>(N, c0 + strip_cols);
                for (int64_t c = c0; c < c1; c++)
                    if (row[c] != 0.0f) { h.ent.push_back(strip_entry(row[c], (uint32_t)(c - c0))); h.row_nnz[(size_t)j]++; h.nnz++; }
                while (h.ent.size() % kStripPad) h.ent.push_back(0);       // all-zero pads: 32-byte segments
                h.row_groups[(size_t)j] += (int32_t)((h.ent.size() - h.soff[seg]) / kStripPad);
            }
        }
    if (h.ent.size() >= (size_t)UINT32_MAX) return SPMV_ERR_UNSUPPORTED;
    h.soff[seg] = (uint32_t)h.ent.size();
    return SPMV_OK;
}

// From CSR(A^T): per band a counting sort of its entries by (row, strip); columns are visited in
// ascending order, so the entries of a segment come out in ascending column order.  Bands are
// independent: a small pool of host threads takes them (same bytes for any thread count).
int pack_strips_csc(int64_t M, int64_t N, const int64_t *col_ptr, const int32_t *row_idx, const float *values,
                    int strip_cols, HostStrips &h)
{
    int64_t nnz = 0;
    for (int64_t k = col_ptr[0]; k < col_ptr[N]; k++) {
        if (row_idx[k] < 0 || row_idx[k] >= M) return SPMV_ERR_ARG;
        nnz += (values[k] != 0.0f);
    }
    if (nnz >= (int64_t)UINT32_MAX) return SPMV_ERR_UNSUPPORTED;
    if (strip_cols <= 0) strip_cols = choose_strip_cols(M, N, nnz);
    if (strip_cols < 32 || strip_cols > kMaxStripCols || strip_cols % 32) return SPMV_ERR_ARG;
    strips_begin(M, N, strip_cols, h);
    h.nnz = nnz;
    const int64_t band_cols = (int64_t)strip_cols * kStripsPerBand;
    const size_t per_band = (size_t)M * kStripsPerBand;
    const int n_threads = pack_threads(h.bands);
    auto run_pool = [&](auto &&fn) {
        if (n_threads == 1) { fn(0); return; }
        std::vector<std::thread> pool;
        for (int t = 0; t < n_threads; t++) pool.emplace_back(fn, t);
        for (std::thread &th : pool) th.join();
    };
    // pass 1: entries per segment (kept in soff for now), stored (padded) entries per band
    std::vector<int64_t> band_first((size_t)h.bands + 1, 0);
    run_pool([&](int t) {
        for (int b = t; b < h.bands; b += n_threads) {
            uint32_t *so = h.soff.data() + (size_t)b * per_band;
            const int64_t cb = b * band_cols, c1 = std::min<int64_t>(N, cb + band_cols);
            for (int64_t c = cb; c < c1; c++) {
                const int s = (int)((c - cb) / strip_cols);
                for (int64_t k = col_ptr[c]; k < col_ptr[c + 1]; k++)
                    if (values[k] != 0.0f) so[(size_t)row_idx[k] * kStripsPerBand + s]++;
            }
            int64_t stored = 0;
            for (size_t i = 0; i < per_band; i++) stored += (so[i] + kStripPad - 1) / kStripPad * kStripPad;
            band_first[(size_t)b + 1] = stored;
        }
    });
    for (int b = 0; b < h.bands; b++) band_first[(size_t)b + 1] += band_first[(size_t)b];
    const int64_t stored_total = band_first[(size_t)h.bands];
    if (stored_total >= (int64_t)UINT32_MAX) return SPMV_ERR_UNSUPPORTED;
    h.ent.assign((size_t)stored_total, 0);                 // pads stay all-zero
    // pass 2: offsets, then the entries (ascending column inside a segment)
    std::vector<std::vector<int32_t>> rn((size_t)n_threads), rg((size_t)n_threads);
    int dup_rc = SPMV_OK;
    run_pool([&](int t) {
        std::vector<int32_t> &row_cnt = rn[(size_t)t], &row_grp = rg[(size_t)t];
        row_cnt.assign((size_t)M, 0); row_grp.assign((size_t)M, 0);
        std::vector<uint32_t> cur(per_band);
        for (int b = t; b < h.bands; b += n_threads) {
            uint32_t *so = h.soff.data() + (size_t)b * per_band;
            uint32_t run = (uint32_t)band_first[(size_t)b];
            for (size_t i = 0; i < per_band; i++) {
                const uint32_t n = so[i], padded = (n + kStripPad - 1) / kStripPad * kStripPad;
                row_cnt[i / kStripsPerBand] += (int32_t)n;
                row_grp[i / kStripsPerBand] += (int32_t)(padded / kStripPad);
                so[i] = run; cur[i] = run; run += padded;
            }
            const int64_t cb = b * band_cols, c1 = std::min<int64_t>(N, cb + band_cols);
            for (int64_t c = cb; c < c1; c++) {
                const int s = (int)((c - cb) / strip_cols);
                const uint32_t lc = (uint32_t)(c - cb - (int64_t)s * strip_cols);
                for (int64_t k = col_ptr[c]; k < col_ptr[c + 1]; k++)
                    if (values[k] != 0.0f) {
                        const size_t seg = (size_t)row_idx[k] * kStripsPerBand + s;
                        const uint32_t p = cur[seg]++;
                        // a repeated (row, column) pair would put two entries of one window on one accumulator
                        if (p > so[seg] && (uint32_t)(h.ent[p - 1] >> 32) == lc + 1u) dup_rc = SPMV_ERR_ARG;
                        h.ent[p] = strip_entry(values[k], lc);
                    }
            }
        }
    });
    if (dup_rc) return dup_rc;
    h.soff.back() = (uint32_t)stored_total;
    for (int t = 0; t < n_threads; t++)
        for (int64_t r = 0; r < M; r++) { h.row_nnz[(size_t)r] += rn[(size_t)t][(size_t)r]; h.row_groups[(size_t)r] += rg[(size_t)t][(size_t)r]; }
    return SPMV_OK;
}

} // namespace spmv

// ==========================================================================================
// Reference host layouts (C-ABI, CPU only)
// ==========================================================================================
namespace {

template <class T> T *take(std::vector<T> &v, int64_t &n)
{
    n = (int64_t)v.size();
    T *p = (T *)std::malloc(std::max<size_t>(v.size(), 1) * sizeof(T));
    if (p && !v.empty()) std::memcpy(p, v.data(), v.size() * sizeof(T));
    return p;
}

inline void set_bit(std::vector<uint32_t> &bm, size_t bit) { bm[bit >> 5] |= 1u << (bit & 31); }

// matrix_csr.cpp:5-23 — one (value,row) list per output column, row_pointers WITHOUT sentinel.
void ref_csr(int M, int N, const float *A, spmv_ref_packed_t *o)
{
    std::vector<int32_t> ptr((size_t)N), idx;
    std::vector<float> val;
    std::vector<int32_t> cnt((size_t)N, 0);
    for (int j = 0; j < M; j++)
        for (int i = 0; i < N; i++) cnt[i] += (A[(size_t)j * N + i] != 0.0f);
    int32_t run = 0;
    for (int i = 0; i < N; i++) { ptr[i] = run; run += cnt[i]; }
    idx.resize((size_t)run); val.resize((size_t)run);
    std::vector<int32_t> cur(ptr);
    for (int j = 0; j < M; j++)
        for (int i = 0; i < N; i++) {
            float v = A[(size_t)j * N + i];
            if (v != 0.0f) { int32_t p = cur[i]++; idx[p] = j; val[p] = v; }
        }
    o->i32_a = take(ptr, o->n_i32_a);
    o->i32_b = take(idx, o->n_i32_b);
    o->f32 = take(val, o->n_f32);
}

// tcsr.cpp:5-38 — tiles slab-major; inside a tile word = column, bit = row; blk_idx with sentinel.
void ref_tcsr(int M, int N, const float *A, spmv_ref_packed_t *o)
{
    const int TR = M / 32, TC = N / 32;
    std::vector<uint32_t> bm((size_t)M * N / 32, 0u);
    std::vector<int32_t> blk((size_t)TR * TC + 1, 0);
    std::vector<float> val;
    size_t tile = 0;
    for (int tc = 0; tc < TC; tc++)
        for (int tr = 0; tr < TR; tr++, tile++) {
            for (int c = 0; c < 32; c++) {
                uint32_t word = 0;
                for (int r = 0; r < 32; r++) {
                    float v = A[(size_t)(tr * 32 + r) * N + tc * 32 + c];
                    if (v != 0.0f) { word |= 1u << r; val.push_back(v); }
                }
                bm[tile * 32 + c] = word;
            }
            blk[tile + 1] = (int32_t)val.size();
        }
    o->i32_a = take(blk, o->n_i32_a);
    o->u32 = take(bm, o->n_u32);
    o->f32 = take(val, o->n_f32);
}

// wsp.cpp:3-40 — column-major bit order (bit i*M+j), ELL-padded values (nz_max_m per column).
void ref_wsp(int M, int N, const float *A, spmv_ref_packed_t *o)
{
    std::vector<uint32_t> bm((size_t)M * N / 32, 0u);
    std::vector<int32_t> cnt((size_t)N, 0);
    for (int j = 0; j < M; j++)
        for (int i = 0; i < N; i++)
            if (A[(size_t)j * N + i] != 0.0f) { cnt[i]++; set_bit(bm, (size_t)i * M + j); }
    int nzmax = 0;
    for (int i = 0; i < N; i++) nzmax = std::max(nzmax, cnt[i]);
    std::vector<float> val((size_t)N * nzmax, 0.0f);
    std::fill(cnt.begin(), cnt.end(), 0);
    for (int j = 0; j < M; j++)
        for (int i = 0; i < N; i++) {
            float v = A[(size_t)j * N + i];
            if (v != 0.0f) val[(size_t)i * nzmax + cnt[i]++] = v;
        }
    o->u32 = take(bm, o->n_u32);
    o->f32 = take(val, o->n_f32);
    o->aux[0] = nzmax; o->aux[1] = N;
}

// asp.cpp:3-14 — dense, 32x32 tiles slab-major, row-major inside the tile.
void ref_asp(int M, int N, const float *A, spmv_ref_packed_t *o)
{
    std::vector<float> val((size_t)M * N);
    const int TR = M / 32;
    for (int j = 0; j < M; j++)
        for (int i = 0; i < N; i++) {
            size_t tile = (size_t)(i / 32) * TR + j / 32;
            val[tile * 1024 + (size_t)(j % 32) * 32 + i % 32] = A[(size_t)j * N + i];
        }
    o->f32 = take(val, o->n_f32);
}

// awsp.cpp:3-49 — tiles slab-major; word = row, bit = column; each tile padded to nz_bk_max.
void ref_awsp(int M, int N, const float *A, spmv_ref_packed_t *o)
{
    const int TR = M / 32, TC = N / 32;
    std::vector<uint32_t> bm((size_t)M * N / 32, 0u);
    std::vector<int32_t> tcnt((size_t)TR * TC, 0);
    for (int j = 0; j < M; j++)
        for (int i = 0; i < N; i++)
            if (A[(size_t)j * N + i] != 0.0f) {
                size_t tile = (size_t)(i / 32) * TR + j / 32;
                tcnt[tile]++;
                bm[tile * 32 + j % 32] |= 1u << (i % 32);
            }
    int bkmax = 0;
    for (int32_t c : tcnt) bkmax = std::max(bkmax, c);
    std::vector<float> val((size_t)TR * TC * bkmax, 0.0f);
    std::fill(tcnt.begin(), tcnt.end(), 0);
    for (int j = 0; j < M; j++)       // row-major sweep = row-major order inside every tile
        for (int i = 0; i < N; i++) {
            float v = A[(size_t)j * N + i];
            if (v != 0.0f) {
                size_t tile = (size_t)(i / 32) * TR + j / 32;
                val[tile * bkmax + tcnt[tile]++] = v;
            }
        }
    o->u32 = take(bm, o->n_u32);
    o->f32 = take(val, o->n_f32);
    o->aux[0] = bkmax;
}

// awsp_ref.cpp:4-58 — bitmap word slab*M+row, bit = column; values per (slab, quarter of M),
// quarter q padded to its max over slabs; warp_nz_offset = inclusive prefix of the maxima.
void ref_awsp_ref(int M, int N, const float *A, spmv_ref_packed_t *o)
{
    const int TC = N / 32, Q = M / 4;
    std::vector<uint32_t> bm((size_t)M * N / 32, 0u);
    std::vector<int32_t> qcnt((size_t)TC * 4, 0);
    for (int j = 0; j < M; j++)
        for (int i = 0; i < N; i++)
            if (A[(size_t)j * N + i] != 0.0f) {
                qcnt[(size_t)(i / 32) * 4 + j / Q]++;
                bm[(size_t)(i / 32) * M + j] |= 1u << (i % 32);
            }
    std::vector<int32_t> off(4, 0);
    int run = 0;
    for (int q = 0; q < 4; q++) {
        int mx = 0;
        for (int s = 0; s < TC; s++) mx = std::max(mx, qcnt[(size_t)s * 4 + q]);
        run += mx; off[q] = run;
    }
    const int stride = off[3];
    std::vector<float> val((size_t)TC * stride, 0.0f);
    std::fill(qcnt.begin(), qcnt.end(), 0);
    for (int j = 0; j < M; j++)
        for (int i = 0; i < N; i++) {
            float v = A[(size_t)j * N + i];
            if (v != 0.0f) {
                int s = i / 32, q = j / Q;
                int base = q ? off[q - 1] : 0;
                val[(size_t)s * stride + base + qcnt[(size_t)s * 4 + q]++] = v;
            }
        }
    o->i32_a = take(off, o->n_i32_a);
    o->u32 = take(bm, o->n_u32);
    o->f32 = take(val, o->n_f32);
}
} // namespace

extern "C" int spmv_ref_pack(int layout, int M, int N, const float *A, spmv_ref_packed_t *out)
{
    if (!out || (!A && (int64_t)M * N > 0)) return spmv::set_error(SPMV_ERR_ARG, "spmv_ref_pack: null argument");
    std::memset(out, 0, sizeof *out);
    if (M < 0 || N < 0 || M % 32 || N % 32)
        return spmv::set_error(SPMV_ERR_SHAPE, "spmv_ref_pack: M and N must be multiples of 32 (tester.cpp:9-10)");
    if ((int64_t)M * N >= ((int64_t)1 << 31))
        return spmv::set_error(SPMV_ERR_SHAPE, "spmv_ref_pack: M*N overflows the reference's int indexing");
    try {
        switch (layout) {
        case SPMV_LAYOUT_CSR: ref_csr(M, N, A, out); break;
        case SPMV_LAYOUT_TCSR: ref_tcsr(M, N, A, out); break;
        case SPMV_LAYOUT_WSP: ref_wsp(M, N, A, out); break;
        case SPMV_LAYOUT_ASP: ref_asp(M, N, A, out); break;
        case SPMV_LAYOUT_AWSP: ref_awsp(M, N, A, out); break;
        case SPMV_LAYOUT_AWSP_REF: ref_awsp_ref(M, N, A, out); break;
        default: return spmv::set_error(SPMV_ERR_ARG, "spmv_ref_pack: unknown layout");
        }
    } catch (const std::bad_alloc &) {
        spmv_ref_packed_free(out);
        return spmv::set_error(SPMV_ERR_NOMEM, "spmv_ref_pack: out of host memory");
    }
    return SPMV_OK;
}

extern "C" void spmv_ref_packed_free(spmv_ref_packed_t *p)
{
    if (!p) return;
    std::free(p->i32_a); std::free(p->i32_b); std::free(p->u32); std::free(p->f32);
    std::memset(p, 0, sizeof *p);
}

// ==========================================================================================
// Device-format inspection (host only)
// ==========================================================================================
namespace {
template <class T> T *dup_vec(const std::vector<T> &v)
{
    T *p = (T *)std::malloc(std::max<size_t>(v.size(), 1) * sizeof(T));
    if (p && !v.empty()) std::memcpy(p, v.data(), v.size() * sizeof(T));
    return p;
}

void dump_wsp(const spmv::HostWsp &w, spmv_packed_dump_t *o)
{
    o->variant = SPMV_WSP; o->index_bits = w.index_bits; o->M = w.M; o->N = w.N; o->nnz = w.nnz; o->groups = w.groups;
    o->slabs = w.panels; o->slab_cols = (int32_t)w.panel_rows;      // wsp: row panels, rows per panel
    o->vals = dup_vec(w.vals); o->n_vals = (int64_t)w.vals.size();
    if (w.index_bits == 16) { o->idx = dup_vec(w.idx16); o->idx_bytes = (int64_t)w.idx16.size() * 2; }
    else { o->idx = dup_vec(w.idx32); o->idx_bytes = (int64_t)w.idx32.size() * 4; }
    o->off = dup_vec(w.colptr); o->n_off = (int64_t)w.colptr.size();
}

void dump_panel(const spmv::HostPanel &h, int variant, spmv_packed_dump_t *o)
{
    o->variant = variant; o->index_bits = h.index_bits; o->slab_cols = h.slab_cols; o->slabs = h.slabs;
    o->row_blocks = h.row_blocks; o->M = h.M; o->N = h.N; o->nnz = h.nnz; o->groups = h.groups;
    o->block_rows = h.block_rows;
    o->vals = dup_vec(h.vals); o->n_vals = (int64_t)h.vals.size();
    if (h.index_bits == 8) { o->idx = dup_vec(h.idx8); o->idx_bytes = (int64_t)h.idx8.size(); }
    else { o->idx = dup_vec(h.idx16); o->idx_bytes = (int64_t)h.idx16.size() * 2; }
    o->off = dup_vec(h.off); o->n_off = (int64_t)h.off.size();
    if (h.tiled) { o->rel = dup_vec(h.rel); o->n_rel = (int64_t)h.rel.size(); }
}

// row strips: vals / idx (u32 column inside the strip) per entry, off = soff, slab_cols = columns per
// strip, slabs = bands, row_blocks = strips per band, block_rows = -1 marks the form
void dump_strips(const spmv::HostStrips &h, spmv_packed_dump_t *o)
{
    o->variant = SPMV_AWSP; o->index_bits = 32; o->slab_cols = h.strip_cols; o->slabs = h.bands;
    o->row_blocks = spmv::kStripsPerBand; o->block_rows = -1; o->M = h.M; o->N = h.N; o->nnz = h.nnz; o->groups = 0;
    std::vector<float> v(h.ent.size());
    std::vector<uint32_t> c(h.ent.size());
    for (size_t k = 0; k < h.ent.size(); k++) {
        const uint32_t bits = (uint32_t)h.ent[k];
        std::memcpy(&v[k], &bits, 4);
        c[k] = (uint32_t)(h.ent[k] >> 32) - 1u;               // the dump shows plain column numbers (pads: 0xffffffff)
    }
    o->vals = dup_vec(v); o->n_vals = (int64_t)v.size();
    o->idx = dup_vec(c); o->idx_bytes = (int64_t)c.size() * 4;
    o->off = dup_vec(h.soff); o->n_off = (int64_t)h.soff.size();
}

int dump_common_check(int variant, int64_t M, int64_t N, spmv_packed_dump_t *out)
{
    if (!out) return spmv::set_error(SPMV_ERR_ARG, "null output");
    std::memset(out, 0, sizeof *out);
    if (variant == SPMV_ASP) return spmv::set_error(SPMV_ERR_UNSUPPORTED, "asp keeps A dense: nothing to dump");
    if (variant < SPMV_WSP || variant > SPMV_TCSR) return spmv::set_error(SPMV_ERR_ARG, "unknown variant %d", variant);
    if (M < 0 || N < 0 || N % 32) return spmv::set_error(SPMV_ERR_SHAPE, "bad shape %lld x %lld", (long long)M, (long long)N);
    return SPMV_OK;
}
} // namespace

extern "C" int spmv_pack_dump_dense(int variant, int64_t M, int64_t N, const float *A, int64_t lda,
                                    const spmv_options_t *opts, spmv_packed_dump_t *out)
{
    int rc = dump_common_check(variant, M, N, out);
    if (rc) return rc;
    if ((!A && M * N > 0) || lda < N) return spmv::set_error(SPMV_ERR_ARG, "bad A / lda");
    try {
        if (variant == SPMV_WSP) {
            spmv::HostWsp w;
            rc = spmv::pack_wsp_dense(M, N, A, lda, opts ? opts->index_bits : 0, w);
            if (!rc) dump_wsp(w, out);
        } else if (opts && opts->chunk_mode == 4 && variant == SPMV_AWSP) {
            spmv::HostStrips h;
            rc = spmv::pack_strips_dense(M, N, A, lda, opts->slab_cols, h);
            if (!rc) dump_strips(h, out);
        } else {
            spmv::HostPanel h;
            rc = spmv::pack_panel_dense(M, N, A, lda, variant == SPMV_TCSR, opts ? opts->slab_cols : 0, h,
                                        opts && opts->chunk_mode == 3);
            if (!rc) dump_panel(h, variant, out);
        }
    } catch (const std::bad_alloc &) { rc = SPMV_ERR_NOMEM; }
    if (rc) return spmv::set_error(rc, "pack failed");
    return SPMV_OK;
}

extern "C" int spmv_pack_dump_csc(int variant, int64_t M, int64_t N, const int64_t *col_ptr,
                                  const int32_t *row_idx, const float *values,
                                  const spmv_options_t *opts, spmv_packed_dump_t *out)
{
    int rc = dump_common_check(variant, M, N, out);
    if (rc) return rc;
    if (!col_ptr) return spmv::set_error(SPMV_ERR_ARG, "col_ptr is null");
    if (spmv::check_csc(M, N, col_ptr, row_idx))
        return spmv::set_error(SPMV_ERR_ARG, "CSR(A^T) input: rows of a column must be in range and strictly ascending (no repeated entries)");
    try {
        if (variant == SPMV_WSP) {
            spmv::HostWsp w;
            rc = spmv::pack_wsp_csc(M, N, col_ptr, row_idx, values, opts ? opts->index_bits : 0, w);
            if (!rc) dump_wsp(w, out);
        } else if (opts && opts->chunk_mode == 4 && variant == SPMV_AWSP) {
            spmv::HostStrips h;
            rc = spmv::pack_strips_csc(M, N, col_ptr, row_idx, values, opts->slab_cols, h);
            if (!rc) dump_strips(h, out);
        } else {
            spmv::HostPanel h;
            rc = spmv::pack_panel_csc(M, N, col_ptr, row_idx, values, variant == SPMV_TCSR, opts ? opts->slab_cols : 0, h,
                                      opts && opts->chunk_mode == 3);
            if (!rc) dump_panel(h, variant, out);
        }
    } catch (const std::bad_alloc &) { rc = SPMV_ERR_NOMEM; }
    if (rc) return spmv::set_error(rc, "pack failed");
    return SPMV_OK;
}

extern "C" void spmv_pack_dump_free(spmv_packed_dump_t *d)
{
    if (!d) return;
    std::free(d->vals); std::free(d->idx); std::free(d->off); std::free(d->rel);
    std::memset(d, 0, sizeof *d);
}
