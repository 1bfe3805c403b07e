// plan.hpp — the opaque spmv_plan and the internal interfaces between capi.cu, the packers
// and the kernel launchers.
#pragma once
#include <cstdint>
#include <vector>

#include <cuda_runtime.h>

#include "formats.hpp"
#include "spmv_b200.h"

namespace spmv {

int set_error(int code, const char *fmt, ...);   // stores the thread's message, returns code
int cuda_error(cudaError_t e, const char *what); // SPMV_ERR_CUDA + message (never exits)

#define SPMV_CUDA(call)                                                        \
    do {                                                                       \
        cudaError_t e_ = (call);                                               \
        if (e_ != cudaSuccess) return ::spmv::cuda_error(e_, #call);           \
    } while (0)

// Programmatic dependent launch.  Every SGEMV kernel is launched with the programmatic stream
// serialization attribute and executes `griddepcontrol.wait` (common.cuh: pdl_wait) before it reads
// or writes global memory that another kernel of the stream may touch (static plan data may be read
// or prefetched before it).  The streaming kernels do not trigger their dependents early, so the next
// kernel of a stream (or captured graph) is placed when the previous grid's CTAs have all exited, and its
// launch and set-up overlap the previous grid's memory flush; after the wait, ordinary stream order holds.
// The short tail kernels (panel_rs_reduce_kernel, strips_reduce_kernel, mg_join_kernel) DO release their
// dependents at entry (pdl_trigger): the next call's first kernel becomes resident, zeroes its accumulators
// and fetches static metadata while the tail kernel runs, then waits for it.
// Measured on one box, graph replays: wsp c2 23.03 -> 22.42 us, awsp c2 20.53 -> 20.01, asp c2
// 24.49 -> 23.99, tcsr c3 11.22 -> 10.66.  An early trigger (griddepcontrol.launch_dependents at
// kernel entry) lets the next grid's CTAs take free slots next to the running ones and unbalances
// the SMs: asp c2 32.8 us.  SPMV_PDL=0 turns the attribute off (the wait is a no-op then).
bool pdl_enabled();
#ifdef __CUDACC__
template <class... P, class... A>
inline cudaError_t launch_k(void (*k)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, A &&...a)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, k, static_cast<P>(a)...);
}
#endif

// host packers (pack_host.cpp)
int pack_wsp_dense(int64_t M, int64_t N, const float *A, int64_t lda, int index_bits, HostWsp &w);
int pack_wsp_csc(int64_t M, int64_t N, const int64_t *col_ptr, const int32_t *row_idx,
                 const float *values, int index_bits, HostWsp &w);
int pack_panel_dense(int64_t M, int64_t N, const float *A, int64_t lda, bool tiled, int slab_cols,
                     HostPanel &P, bool lane_owned = false);
int pack_panel_csc(int64_t M, int64_t N, const int64_t *col_ptr, const int32_t *row_idx,
                   const float *values, bool tiled, int slab_cols, HostPanel &P, bool lane_owned = false);
int choose_slab_cols(int64_t M, int64_t N, int64_t nnz);
int check_csc(int64_t M, int64_t N, const int64_t *col_ptr, const int32_t *row_idx);   // SPMV_OK or SPMV_ERR_ARG
int pack_strips_dense(int64_t M, int64_t N, const float *A, int64_t lda, int strip_cols, HostStrips &h);
int pack_strips_csc(int64_t M, int64_t N, const int64_t *col_ptr, const int32_t *row_idx, const float *values,
                    int strip_cols, HostStrips &h);
int choose_strip_cols(int64_t M, int64_t N, int64_t nnz);
void wsp_choose_panels(HostWsp &w, int64_t nnz, int index_bits_opt);   // sets panels, panel_rows, index_bits

struct DevWsp {
    uint32_t *colptr = nullptr;
    float *vals = nullptr;
    void *idx = nullptr;
    int index_bits = 16;
    int warps_per_col = 1;
    bool x_in_smem = true;
    int panels = 1;             // row panels (tall matrices)
    int64_t panel_rows = 0;
};

struct DevAsp {
    float *A = nullptr;     // dense row-major, leading dimension ld (multiple of 4)
    int64_t ld = 0;
    int tile_cols = 0;      // columns per CTA
    int rows_per_split = 0;
    // TMA row-gather path (asp.cu: asp_tma_kernel): a 2-D tensor map of A with a (256 columns x 1 row) box, encoded at
    // the first launch for the buffer it describes (a clone has its own A, so it encodes its own)
    alignas(64) unsigned char tmap[128] = {};
    const float *tmap_for = nullptr;
    int tma_state = 0;      // 0: not tried yet, 1: usable, -1: not available (driver entry point, alignment)
};

struct DevPanel {
    uint32_t *off = nullptr;
    uint16_t *rel = nullptr;
    float *vals = nullptr;
    void *idx = nullptr;
    int slab_cols = 256;
    int index_bits = 8;
    int slabs = 0, row_blocks = 0;
    int warps = 8;
    int kmax = 2;               // partial rows reserved per CTA (pieces of its flat range)
    bool tiled = false;
    bool multirow = false;      // several short rows per 32-group chunk
    int block_rows = 0;         // > 0: lane-owned blocks (formats.hpp), off is per (slab, block)
    int lob_blocks = 0;
    bool batch_ok = false;      // scratch holds two vectors' partial rows (spmv_run_batch on awsp / tcsr)
    // register-staged form (panel_rs.cu) for one-row-per-chunk plans: rs_grid > 0 when spmv_run takes it
    int rs_grid = 0, rs_kmax = 0, rs_smem = 0;
    bool rs_by_block = false;   // pieces of >= 256 rows: a warp owns whole 32-row blocks
    float *rs_partial = nullptr;
};

// row strips (formats.hpp: HostStrips; strips.cu)
struct DevStrips {
    uint2 *ent = nullptr;
    uint32_t *soff = nullptr;
    int strip_cols = 0;         // > 0: the plan is in the row-strip form
    int bands = 0;
    int ctas_per_band = 1;
};

} // namespace spmv

struct spmv_plan {
    int variant = 0;
    int device = 0;
    int64_t M = 0, N = 0, nnz = 0;
    int sm_count = 148;
    int max_smem_optin = 0;

    spmv::DevWsp wsp;
    void *bufs = nullptr;       // spmv::PlanBufs (capi.cu): registry of device allocations
    void *wsp_state = nullptr;  // spmv::WspState (wsp.cu): the length bins
    int wsp_team = 0;
    spmv::DevAsp asp;
    spmv::DevPanel panel;
    spmv::DevStrips strips;

    // split-reduction scratch (asp/awsp/tcsr)
    int row_splits = 1;
    int col_tiles = 0;          // number of output tiles (tickets)
    int tile_width = 0;
    float *partial = nullptr;   // [row_splits][col_tiles*tile_width]
    unsigned *tickets = nullptr;

    // launch geometry of the main kernel
    dim3 grid{1, 1, 1};
    int block = 0;
    int smem = 0;
    int kernels_per_run = 1;

    // host-side accounting
    int64_t device_bytes = 0, scratch_bytes = 0;
    std::vector<int32_t> row_nnz;    // awsp/tcsr: stored nnz per row
    std::vector<int32_t> row_groups; //            groups per row
    std::vector<int32_t> row_segs;   //            non-empty segments per row
    int64_t fmt_groups = 0;          // wsp / panel: total groups
    int64_t off_bytes = 0;           // bytes of the offset tables

    // staging buffers for spmv_run_host
    float *d_x = nullptr, *d_y = nullptr;
    cudaStream_t stream = nullptr;  // owned, used by spmv_run_host
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // spmv_run_host with the same pinned (x, y) pair again: H2D + kernels + D2H replayed as one graph
    const float *graph_x = nullptr;
    float *graph_y = nullptr;
    cudaGraphExec_t graph_exec = nullptr;
};

namespace spmv {
// kernel launchers (one per .cu); all are asynchronous on `st`
struct YDst;   // common.cuh: where y goes (one pointer, or every rank's copy in the sharded case)
int launch_wsp(spmv_plan *p, const float *d_x, const YDst &yd, cudaStream_t st);
int launch_wsp_batch(spmv_plan *p, const float *d_x, long long ldx, const YDst &yd, long long ldy, int B, cudaStream_t st);
int launch_asp(spmv_plan *p, const float *d_x, const YDst &yd, cudaStream_t st);
int launch_asp_batch(spmv_plan *p, const float *d_x, long long ldx, const YDst &yd, long long ldy, int B, cudaStream_t st);
int launch_panel(spmv_plan *p, const float *d_x, const YDst &yd, cudaStream_t st);
int launch_panel_rs(spmv_plan *p, const float *d_x, const YDst &yd, cudaStream_t st);
int configure_panel_rs(spmv_plan *p, const HostPanel &h);
int launch_panel_batch(spmv_plan *p, const float *d_x, long long ldx, const YDst &yd, long long ldy, int B, cudaStream_t st);
int launch_strips(spmv_plan *p, const float *d_x, const YDst &yd, cudaStream_t st);
int launch_compact(const float *d_x, int64_t M, int32_t *d_idx, float *d_val, int32_t *d_count,
                   void *d_scratch, size_t scratch_bytes, cudaStream_t st);
size_t compact_scratch_bytes(int64_t M);
// geometry setup (fills grid/block/smem, allocates scratch)
int configure_wsp(spmv_plan *p, const HostWsp &w, const spmv_options_t *o);
int configure_asp(spmv_plan *p, const spmv_options_t *o);
int configure_panel(spmv_plan *p, const HostPanel &h, const spmv_options_t *o);
int configure_strips(spmv_plan *p, const HostStrips &h, const spmv_options_t *o);
void destroy_wsp_state(spmv_plan *p);
int clone_wsp_state(const spmv_plan *src, spmv_plan *dst);
int alloc_split_scratch(spmv_plan *p, int copies = 1);   // partial + tickets from row_splits/col_tiles/tile_width (x copies)
int alloc_panel_scratch(spmv_plan *p, size_t partial_floats, size_t tickets);
// a device allocation owned by the plan; `slot` must be a pointer member of *p (capi.cu)
int plan_alloc(spmv_plan *p, void **slot, size_t bytes, bool zero);
// device packers (pack_dev.cu): dense matrix in HBM -> the plan's device arrays, bit-identical to
// the host packers; the Host* argument receives only what the geometry setup needs
int pack_wsp_device(spmv_plan *p, const float *d_A, int64_t lda, int index_bits_opt, HostWsp &w);
int pack_panel_device(spmv_plan *p, const float *d_A, int64_t lda, bool tiled, int slab_cols_opt, HostPanel &h);
// row strips from CSR(A^T) in device memory (bit-identical to pack_strips_csc)
int pack_strips_csc_device(spmv_plan *p, const int64_t *d_col_ptr, const int32_t *d_row_idx, const float *d_values,
                           int strip_cols_opt, HostStrips &h);
} // namespace spmv
