/*
 * spmv_b200.h — C-ABI of the B200-native sparse SGEMV library (libspmv_b200.so).
 *
 * Operation (reference README.md:29-35, tester.cpp:36-45):
 *     Y = X * A,  X: float[M],  A: float[M][N] row-major (A[j*N+i]),  Y: float[N]
 *     y[i] = sum_j x[j] * A[j][i]
 *
 * The reference (PACTHEMAN123/spMV-test) has no FFI layer: its boundary is the
 * C++ launcher set in src/include/kernel.hpp:8-17, each of which packs a dense
 * host matrix, uploads, launches once, downloads and frees.  This header is the
 * plan/handle API that sits underneath drop-in re-implementations of those
 * launchers (spmv_test_b200/host/launchers.cpp) and that tests / bench.py bind
 * with ctypes.  Plain pointers and sizes only; no C++ or torch types.
 *
 * Conventions
 *   - every function returns 0 (SPMV_OK) or a negative spmv_status_t;
 *     spmv_last_error() gives the message of the calling thread's last failure.
 *     Nothing below the C++ shim calls exit() (the reference's CUDA_CHECK,
 *     kernel.hpp:21-28, does; the shim keeps that behaviour).
 *   - "d_" pointers are device pointers on the plan's device, "h_"/unprefixed
 *     pointers are host pointers.  stream is a cudaStream_t passed as void*.
 *   - spmv_run() is asynchronous, allocation-free and CUDA-graph capturable.
 *   - results are deterministic: no floating-point atomics anywhere; a plan reproduces its
 *     results bit for bit (the work decomposition depends only on the shape and the device).
 *   - a plan owns its split-reduction scratch: do not run ONE plan on two streams at the same
 *     time (calls on one stream may overlap freely with other plans); spmv_plan_clone() gives an
 *     independent copy.
 *   - there is no CPU fallback: if no CUDA device is usable the call fails.
 */
#ifndef SPMV_B200_H
#define SPMV_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPMV_B200_ABI_VERSION 1

#if defined(__GNUC__)
#define SPMV_API __attribute__((visibility("default")))
#else
#define SPMV_API
#endif

typedef enum spmv_status {
    SPMV_OK            =  0,
    SPMV_ERR_ARG       = -1,  /* null pointer, bad enum, bad option            */
    SPMV_ERR_SHAPE     = -2,  /* shape precondition violated (see below)       */
    SPMV_ERR_CUDA      = -3,  /* a CUDA runtime call failed                    */
    SPMV_ERR_NOMEM     = -4,  /* host allocation failed                        */
    SPMV_ERR_UNSUPPORTED = -5 /* valid request this build cannot serve         */
} spmv_status_t;

/*
 * Variants.  Each replaces one launcher family of the reference:
 *   SPMV_WSP   weight-sparse       wsp_gemv_gpu / csr_naive_gemv_gpu   (kernel.hpp:11,13; wsp.cu:4-138; csr_naive.cu:6-23)
 *   SPMV_ASP   activation-sparse   asp_gemv_gpu                        (kernel.hpp:14; asp.cu:6-211)
 *   SPMV_AWSP  both                awsp_gemv_gpu / awsp_ref_gemv_gpu   (kernel.hpp:15-16; awsp.cu:5-317; awsp_ref.cu:6-185)
 *   SPMV_TCSR  tiled bitmap-CSR    csr_tiling_gemv_gpu                 (kernel.hpp:12; csr_tiling.cu:24-114)
 */
typedef enum spmv_variant {
    SPMV_WSP  = 0,
    SPMV_ASP  = 1,
    SPMV_AWSP = 2,
    SPMV_TCSR = 3
} spmv_variant_t;

/* Tuning knobs; zero-initialise for defaults.  struct_size must be sizeof(spmv_options_t). */
typedef struct spmv_options {
    uint32_t struct_size;
    int32_t  row_splits;     /* asp: CTAs along M per column tile; awsp/tcsr: CTAs per slab (0 = auto) */
    int32_t  warps_per_col;  /* wsp: warps cooperating on one column, 1/2/4/8; awsp/tcsr: warps per CTA (0 = auto) */
    int32_t  index_bits;     /* wsp: 16 or 32 bit row indices (0 = auto: 16 if M<65536)  */
    int32_t  slab_cols;      /* awsp/tcsr: columns per slab, power of two 256..4096 (0 = auto from density) */
    int32_t  chunk_mode;     /* awsp/tcsr: 0 = auto (spmv_plan_create_csc: awsp matrices under ~1.2 % density take mode 4),
                                1 = one row per 32-group chunk, 2 = short rows packed into shared chunks,
                                3 = lane-owned blocks: for very sparse matrices (segments of a few non-zeros) whose
                                activations are mostly non-zero — every stored non-zero is read, x is a multiplier,
                                and a chunk retires in one pass (host packer only; slab_cols 1024..4096, default 2048),
                                4 = row strips (awsp only): very sparse matrices (about 1 % dense) — 8-byte entries,
                                one row segment per 32-lane window, rows with x == 0 are never read; slab_cols then
                                means columns per strip (multiple of 32, 32..2112; 0 = about 20 non-zeros per row
                                and strip); host packer only */
    int32_t  pack_mode;      /* dense input: 0 = auto, 1 = pack on the host, 2 = pack on the GPU (the dense matrix is
                                staged in HBM first; needs M*N*4 bytes of spare device memory).  Both give the
                                same bytes. */
    int32_t  reserved[1];
} spmv_options_t;

typedef struct spmv_plan spmv_plan_t;   /* opaque */

typedef struct spmv_plan_info {
    int32_t  variant;
    int64_t  M, N;
    int64_t  nnz;               /* stored non-zeros (A != 0.0f); M*N for asp                */
    int64_t  device_bytes;      /* bytes of the packed format resident in HBM               */
    int64_t  scratch_bytes;     /* split partial sums + counters                             */
    int32_t  kernels_per_run;   /* kernel launches one spmv_run() issues                     */
    int32_t  grid_x, grid_y, block; /* launch geometry of the main kernel                    */
    int32_t  smem_bytes;        /* dynamic shared memory of the main kernel                  */
    int32_t  index_bits;        /* wsp only                                                  */
    int32_t  row_splits;
    int32_t  warps_per_col;
    int32_t  slab_cols;         /* awsp/tcsr                                                 */
    int32_t  reserved[3];
} spmv_plan_info_t;

/* ---- library ----------------------------------------------------------------------------- */
SPMV_API int         spmv_abi_version(void);
SPMV_API const char *spmv_last_error(void);
/* Number of CUDA devices visible, or a negative status.  Never throws, never exits. */
SPMV_API int         spmv_device_count(void);
/* Plans and groups live on the calling thread's current CUDA device (the reference uses device 0
 * of CUDA_VISIBLE_DEVICES throughout); this selects it for callers without the CUDA runtime headers. */
SPMV_API int         spmv_set_device(int device);

/* ---- plans ------------------------------------------------------------------------------- */
/*
 * Pack a dense row-major host matrix (leading dimension lda >= N, so a column
 * slab of a wider matrix can be passed as A + col_begin with lda = N_total —
 * the multi-GPU partitioner does exactly that) into the variant's device format
 * and upload it to the current CUDA device.  Replaces the pack+cudaMalloc+H2D
 * prologue every reference launcher repeats (e.g. awsp.cu:323-344).
 *
 * Preconditions (reference tester.cpp:9-10 asserts M%32==0 && N%32==0; its kernels
 * silently need more): here M >= 0 is arbitrary, N % 32 == 0 is
 * required and anything else is rejected with SPMV_ERR_SHAPE.  M*N == 0 is legal.
 * A weight is "zero" iff value == 0.0f is true (so -0.0f is zero, NaN is kept):
 * matrix_csr.cpp:15, wsp.cpp:17, awsp.cpp:20.
 */
SPMV_API int spmv_plan_create_dense(int variant, int64_t M, int64_t N, const float *A, int64_t lda,
                           const spmv_options_t *opts, spmv_plan_t **out);

/*
 * The same from a dense row-major matrix that is already in DEVICE memory (current device):
 * count, prefix sum, fill and the in-chunk ordering all run as kernels (the
 * reference's packers are single-threaded host loops, wsp.cpp:25-37, awsp.cpp:30-46, and every
 * launcher call pays for them).  The resulting plan is bit-identical to the one
 * spmv_plan_create_dense builds from the same values.  d_A is only read and can be freed
 * afterwards (asp keeps its own copy).  Synchronous.
 */
SPMV_API int spmv_plan_create_dense_device(int variant, int64_t M, int64_t N, const float *d_A, int64_t lda,
                                  const spmv_options_t *opts, spmv_plan_t **out);

/*
 * Direct-to-sparse construction for shapes whose dense form cannot exist
 * (BASELINE configs 4 and 5).  Input is the reference CSRMatrix orientation
 * (matrix_csr.cpp:8-22: one list per OUTPUT column i, entries (row j, value) with
 * j ascending) but with 64-bit pointers and the N+1 sentinel the reference omits.
 * Inside a column the rows must ascend STRICTLY and lie in [0, M): a repeated (row, column)
 * entry, an unsorted list or an out-of-range row is rejected with SPMV_ERR_ARG (entries are
 * never summed), so every variant sees the same matrix.
 * Supported for SPMV_WSP, SPMV_AWSP and SPMV_TCSR.
 */
SPMV_API int spmv_plan_create_csc(int variant, int64_t M, int64_t N, const int64_t *col_ptr,
                         const int32_t *row_idx, const float *values,
                         const spmv_options_t *opts, spmv_plan_t **out);

/*
 * The same from CSR(A^T) arrays that are already in DEVICE memory (current device): check, count,
 * scan and fill run as kernels, so a config-5-size slab (86 M non-zeros) is packed in milliseconds
 * instead of the host packers' seconds.  Supported for the row-strip form (SPMV_AWSP
 * with opts->chunk_mode == 4); the plan is bit-identical to spmv_plan_create_csc on the same arrays.
 * The input arrays are only read and can be freed afterwards.  Synchronous.
 */
SPMV_API int spmv_plan_create_csc_device(int variant, int64_t M, int64_t N, const int64_t *d_col_ptr,
                                const int32_t *d_row_idx, const float *d_values,
                                const spmv_options_t *opts, spmv_plan_t **out);

SPMV_API int  spmv_plan_info(const spmv_plan_t *plan, spmv_plan_info_t *info);
SPMV_API void spmv_plan_destroy(spmv_plan_t *plan);

/*
 * Device-side copy of a plan (same device): a second, independent resident copy of the packed
 * matrix.  The benchmark rotates over clones whose total size exceeds the L2 so that every
 * timed call streams from HBM; a server would use it to replicate a hot matrix.
 */
SPMV_API int spmv_plan_clone(const spmv_plan_t *plan, spmv_plan_t **out);

/*
 * Plan files (packing is O(M*N) host work the reference repeats on every call,
 * e.g. awsp.cpp:3-49).  spmv_plan_save writes the packed format of a plan; spmv_plan_load uploads
 * it again without re-packing and re-derives the launch geometry for the current device (opts as
 * in spmv_plan_create_*).  A loaded plan computes bit-identical results.  Files are validated on
 * load (sizes, offsets, index ranges); a damaged file fails with SPMV_ERR_ARG.
 */
SPMV_API int spmv_plan_save(const spmv_plan_t *plan, const char *path);
SPMV_API int spmv_plan_load(const char *path, const spmv_options_t *opts, spmv_plan_t **out);

/*
 * Bytes one call moves for a given activation vector (host copy of x):
 *   alg_bytes  — the format-independent figure of SURVEY §8d:
 *                8*nnz_t + 4*(N+1) + 4*M + 4*N   (wsp/tcsr: nnz_t = all stored nnz;
 *                awsp: stored nnz in rows with x[j] != 0),  asp: 4*M_nz*N + 4*M + 4*N
 *   phys_bytes — bytes of the packed arrays the kernel actually has to read for
 *                this x in this library's format, plus x and y.
 */
SPMV_API int spmv_plan_traffic(const spmv_plan_t *plan, const float *x, double *alg_bytes,
                      double *phys_bytes, int64_t *nnz_touched);

/* ---- execution --------------------------------------------------------------------------- */
/* y = x*A on device buffers (x: M floats, y: N floats; 16-byte aligned). */
SPMV_API int spmv_run(spmv_plan_t *plan, const float *d_x, float *d_y, void *stream);

/*
 * The same with an activation fused into the stores of y: y = act(x·A).  In a decode FFN
 * (BASELINE configs 2 -> 3: up-projection, activation, down-projection) the intermediate then
 * leaves the first kernel already rectified, and the second call's fused `x != 0` compaction skips
 * its zeros: two launches, no elementwise kernel, no compaction pass in between.
 * ReLU is `v < 0 ? 0 : v` (NaN is kept).
 */
typedef enum spmv_activation { SPMV_ACT_NONE = 0, SPMV_ACT_RELU = 1 } spmv_activation_t;
SPMV_API int spmv_run_act(spmv_plan_t *plan, const float *d_x, float *d_y, int activation, void *stream);

/*
 * Batched (multi-vector) form, SURVEY section 8f-2: Y[b] = X[b] * A for b < batch, X row-major
 * batch x M (row stride ldx), Y row-major batch x N (row stride ldy, multiple of 4), device
 * pointers.  wsp plans stream A once per group of 4 (or 2) vectors with all of them in shared
 * memory, asp plans stream a row once if any of the 4 (or 2) vectors is active there, awsp / tcsr plans
 * (one row per chunk: the dense-ish shapes) stream a row segment once for 2 vectors if either is active
 * there, so A's bytes are reused; multi-row, lane-owned and row-strip plans run the vectors one by one.
 * wsp / asp (and every vector that runs alone): Y[b] is bit-identical to spmv_run on X[b].  awsp / tcsr
 * pairs: reproducible run to run and inside the same parity gate, but the rows are dealt to the warps by
 * their rank among the rows active in EITHER vector, so the order of the sums — and with it the last
 * bits — can differ from the single-vector call.
 */
SPMV_API int spmv_run_batch(spmv_plan_t *plan, int batch, const float *d_X, int64_t ldx, float *d_Y, int64_t ldy,
                            void *stream);

/*
 * Column-sharded multi-GPU execution with the all-gather fused into the kernel epilogue
 * (the reference is single-GPU).  This rank's plan covers the output columns
 * [offset, offset + N) of the full y; its final stores go to that slice of EVERY rank's copy of
 * y: d_y_dst[k] (k < n_dst <= 8) are the base pointers of the ranks' full-y buffers as mapped
 * into this process (peer-accessible / symmetric memory, e.g. torch symmetric memory
 * buffer_ptrs); if d_y_multicast is non-NULL it is the NVSwitch multicast alias of those buffers
 * and one multimem.st per element replaces the n_dst peer stores.  The caller separates
 * consecutive calls with a cross-rank barrier (and alternates two y buffers), exactly as it
 * would around an all-gather.  With n_dst = 1 and offset = 0 this is spmv_run.
 */
SPMV_API int spmv_run_scatter(spmv_plan_t *plan, const float *d_x, int n_dst, float *const *d_y_dst,
                              float *d_y_multicast, int64_t offset, void *stream);

/*
 * Host-buffer convenience: H2D x, run, D2H y, synchronise — the per-call part of a
 * reference launcher once the matrix is resident (awsp.cu:342-381).  If timing_ms is
 * non-NULL it receives the device time of the kernel(s) alone (the region the
 * reference's TIME_KERNEL macro brackets, kernel.hpp:31-48).
 */
SPMV_API int spmv_run_host(spmv_plan_t *plan, const float *x, float *y, float *timing_ms);

/*
 * Activation compaction (the x != 0.0f test of asp.cu:23, awsp.cu:98,127,228,258,
 * awsp_ref.cu:52,96, made an explicit pass): writes the ascending list of rows j
 * with x[j] != 0.0f, their values, and the count.  d_idx/d_val need M entries.
 * Deterministic and bit-exact with the oracle.
 */
SPMV_API int    spmv_compact_x(const float *d_x, int64_t M, int32_t *d_idx, float *d_val,
                      int32_t *d_count, void *d_scratch, size_t scratch_bytes, void *stream);
/* Device scratch spmv_compact_x needs for this M (0 for M <= 32768: d_scratch may be NULL). */
SPMV_API size_t spmv_compact_x_scratch_bytes(int64_t M);

/* ---- multi-GPU partitioner (host only) --------------------------------------------------- */
/*
 * Column-slab partition of the N outputs over `parts` GPUs (BASELINE config 5; the reference
 * is single-GPU): bounds[g] .. bounds[g+1] is the contiguous range GPU g owns.
 * Boundaries are multiples of `align` (use the slab width, >= 32).  With col_ptr == NULL the
 * split is by column count; with a CSR(A^T) col_ptr (N+1 entries) it is balanced by
 * non-zeros.  Every GPU then runs the single-GPU kernel on its slab and the Y slices are
 * joined by one all-gather; no reduction crosses GPUs, so results stay deterministic.
 */
SPMV_API int spmv_partition_columns(int64_t N, int parts, int64_t align, const int64_t *col_ptr,
                           int64_t *bounds /* parts+1 */);

/* ---- column-sharded groups: several slabs per GPU, several GPUs ------------------------------- */
/*
 * SURVEY section 8e / BASELINE config 5 (the reference is single-GPU, one matrix per launcher call,
 * e.g. awsp.cu:319-388).  A group describes one rank's share of a column-sharded matrix: its
 * plans (each a slab of columns at a column offset of the full y) and the shared block that holds
 * two alternating copies of the full y plus the arrival flags.  spmv_mg_run launches the rank's
 * plans back to back; their epilogues store into EVERY rank's y (spmv_run_scatter), and a one-warp
 * arrival kernel replaces the all-gather's synchronisation: it publishes this rank's epoch to all
 * peers and waits for theirs on the device (no host round trip; CUDA-graph capturable — capture an
 * even number of calls, the y buffers alternate).  With world == 1 it is simply "several slabs,
 * one y".  No reduction crosses ranks: results are bit-identical to running the slabs alone.
 *
 *   one process per GPU    spmv_mg_create(NULL block) on every rank, exchange spmv_mg_ipc_handle()
 *                          bytes through the caller's own transport (torch.distributed, MPI, a
 *                          file), spmv_mg_connect_ipc();  or allocate the blocks as symmetric memory
 *                          (spmv_mg_block_bytes), pass this rank's block to spmv_mg_create and all
 *                          ranks' mappings (+ the NVSwitch multicast alias, if any) to
 *                          spmv_mg_connect_ptrs();
 *   one process, n GPUs    spmv_mg_create_group() (peer access), spmv_mg_group_run_host().
 */
typedef struct spmv_mg spmv_mg_t;   /* opaque */
SPMV_API size_t spmv_mg_block_bytes(int64_t N_total);
SPMV_API int  spmv_mg_create(int64_t M, int64_t N_total, int rank, int world, void *local_block, spmv_mg_t **out);
SPMV_API void spmv_mg_destroy(spmv_mg_t *mg);
SPMV_API int  spmv_mg_ipc_handle(spmv_mg_t *mg, void *handle64 /* 64 bytes out */);
SPMV_API int  spmv_mg_connect_ipc(spmv_mg_t *mg, const void *handles /* world x 64 bytes, rank order */);
SPMV_API int  spmv_mg_connect_ptrs(spmv_mg_t *mg, void *const *blocks /* world */, void *multicast_block /* or NULL */);
/* plan: a slab of plan->N columns at col_offset (multiple of 4) of the full y; the caller keeps ownership */
SPMV_API int  spmv_mg_add_plan(spmv_mg_t *mg, spmv_plan_t *plan, int64_t col_offset);
/* asynchronous on `stream`; *d_y = this rank's copy of the full y, complete when the stream reaches this point */
SPMV_API int  spmv_mg_run(spmv_mg_t *mg, const float *d_x, void *stream, const float **d_y);
/* H2D x, spmv_mg_run, D2H y[y_begin, y_begin + y_count), synchronise (y_count = 0: no copy back).  When the range lies inside
 * this rank's own slabs and there are several of them, each slab's columns travel to the host (second stream, one event per
 * slab) while the later slabs still compute; pinned host memory makes that overlap real. */
SPMV_API int  spmv_mg_run_host(spmv_mg_t *mg, const float *x, float *y, int64_t y_begin, int64_t y_count);
/* SPMV_OK, or SPMV_ERR_CUDA if an arrival wait timed out (a peer died); synchronous */
SPMV_API int  spmv_mg_status(spmv_mg_t *mg);
SPMV_API int  spmv_mg_create_group(int64_t M, int64_t N_total, int n_dev, const int *devices, spmv_mg_t **out /* n_dev */);
SPMV_API int  spmv_mg_group_run_host(spmv_mg_t *const *groups, int n_dev, const float *x, float *y /* N_total */);

/* ---- device-format inspection (host only; no GPU needed) ---------------------------------- */
/*
 * Packs like spmv_plan_create_dense / _csc would, but hands the host image of the device
 * format back instead of uploading it: the CPU test-suite decodes it to check the packers
 * (group padding, offsets, column ids) without a GPU.  Not a compute path.
 *   SPMV_WSP        vals[4*(groups+1)], idx (u16|u32)[4*(groups+1)], off = colptr[panels*N+1]
 *                   (slabs = row panels, slab_cols = rows per panel; one list per (panel, column))
 *   SPMV_AWSP       vals, idx (u8|u16)[4*groups], off[slabs*(M+1)]
 *   SPMV_TCSR       vals, idx, off = tile_off[slabs*(row_blocks+1)], rel[slabs*row_blocks*32]
 */
typedef struct spmv_packed_dump {
    int32_t  variant, index_bits, slab_cols, slabs, row_blocks, block_rows /* > 0: lane-owned blocks */;
    int64_t  M, N, nnz, groups;
    float    *vals;  int64_t n_vals;
    void     *idx;   int64_t idx_bytes;
    uint32_t *off;   int64_t n_off;
    uint16_t *rel;   int64_t n_rel;
} spmv_packed_dump_t;

SPMV_API int  spmv_pack_dump_dense(int variant, int64_t M, int64_t N, const float *A, int64_t lda,
                                   const spmv_options_t *opts, spmv_packed_dump_t *out);
SPMV_API int  spmv_pack_dump_csc(int variant, int64_t M, int64_t N, const int64_t *col_ptr,
                                 const int32_t *row_idx, const float *values,
                                 const spmv_options_t *opts, spmv_packed_dump_t *out);
SPMV_API void spmv_pack_dump_free(spmv_packed_dump_t *d);

/* ---- reference host layouts (CPU only; used by the drop-in format classes) ----------------- */
/*
 * Bit-exact re-implementations of the reference's six host packers.  No GPU needed.
 *   layout                 i32_a                 i32_b        u32        f32          aux
 *   SPMV_LAYOUT_CSR        row_pointers[N]       col_idx[nnz] -          values[nnz]  -              matrix_csr.cpp:5-23
 *   SPMV_LAYOUT_TCSR       blk_idx[(M/32)(N/32)+1] -          bitmaps    values[nnz]  -              tcsr.cpp:5-38
 *   SPMV_LAYOUT_WSP        -                     -            bitmaps    values       nz_max_m, nz_max_n   wsp.cpp:3-40
 *   SPMV_LAYOUT_ASP        -                     -            -          values[M*N]  -              asp.cpp:3-14
 *   SPMV_LAYOUT_AWSP       -                     -            bitmaps    values       nz_bk_max      awsp.cpp:3-49
 *   SPMV_LAYOUT_AWSP_REF   warp_nz_offset[4]     -            bitmaps    values       -              awsp_ref.cpp:4-58
 * Shape rule: M % 32 == 0 and N % 32 == 0 (tester.cpp:9-10); AWSP_REF also M % 4 == 0.
 */
typedef enum spmv_layout {
    SPMV_LAYOUT_CSR = 0,
    SPMV_LAYOUT_TCSR = 1,
    SPMV_LAYOUT_WSP = 2,
    SPMV_LAYOUT_ASP = 3,
    SPMV_LAYOUT_AWSP = 4,
    SPMV_LAYOUT_AWSP_REF = 5
} spmv_layout_t;

typedef struct spmv_ref_packed {
    int32_t  *i32_a;  int64_t n_i32_a;
    int32_t  *i32_b;  int64_t n_i32_b;
    uint32_t *u32;    int64_t n_u32;
    float    *f32;    int64_t n_f32;
    int32_t   aux[4];
} spmv_ref_packed_t;

SPMV_API int  spmv_ref_pack(int layout, int M, int N, const float *A, spmv_ref_packed_t *out);
SPMV_API void spmv_ref_packed_free(spmv_ref_packed_t *p);

#ifdef __cplusplus
}
#endif
#endif /* SPMV_B200_H */
