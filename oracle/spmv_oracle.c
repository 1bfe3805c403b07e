/*
 * spmv_oracle.c — CPU ORACLE.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C restatement of the algorithms of PACTHEMAN123/spMV-test for the
 * sparse SGEMV hot path Y = X*A.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this library; the
 * product (libspmv_b200.so) never links, loads or calls it.
 *
 * Parity pin: PINNED.  tests/test_oracle_pinned.py checks every function below
 * bit-for-bit against the reference's own sources compiled in place
 * (oracle/_ref/libspmv_ref_cpu.so, recipe oracle/Makefile) and against the
 * digests in tests/golden/ generated from that build (tests/golden/make_golden.py).
 * The reference itself ships no golden vectors: its only check is the
 * run-time CompareY (tester.cpp:74-88).
 *
 * Every function cites the reference file:line it restates.  Build with
 * -ffp-contract=off so `acc += a*b` stays a separate multiply and add, as in the
 * reference host build (g++ -O2, no -march: no FMA); the *_gpu emulations call
 * fmaf() explicitly where nvcc contracts the reference kernels' a*b+c.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------
 * The oracle of oracles: dense sequential fp32 product.   tester.cpp:36-45 (SgemvCPU)
 * ---------------------------------------------------------------------------------------- */
ORC_API void orc_sgemv_dense(int M, int N, const float *A, const float *x, float *y)
{
    for (int i = 0; i < N; i++) {
        float acc = 0.0f;
        for (int j = 0; j < M; j++)
            acc += x[j] * A[(size_t)j * N + i];
        y[i] = acc;
    }
}

/* Same product in fp64, plus the magnitude sum used to normalise the tolerance
 * (|y - y_ref|_i <= 1e-5 * sum_j |x_j a_ji|).  Not in the reference. */
ORC_API void orc_sgemv_dense_f64(int M, int N, const float *A, const float *x, double *y,
                                 double *abs_sum)
{
    for (int i = 0; i < N; i++) { y[i] = 0.0; if (abs_sum) abs_sum[i] = 0.0; }
    for (int j = 0; j < M; j++) {
        const double xj = x[j];
        if (xj == 0.0) continue;
        const float *row = A + (size_t)j * N;
        for (int i = 0; i < N; i++) {
            const double p = xj * (double)row[i];
            y[i] += p;
            if (abs_sum) abs_sum[i] += fabs(p);
        }
    }
}

/* Activation zero test: x != 0.0f (asp.cu:23; awsp.cu:228,258; awsp_ref.cu:52,96).
 * -0.0f counts as zero, NaN counts as non-zero.  Returns the count. */
ORC_API int orc_compact_x(int M, const float *x, int32_t *idx, float *val)
{
    int n = 0;
    for (int j = 0; j < M; j++)
        if (x[j] != 0.0f) { idx[n] = j; val[n] = x[j]; n++; }
    return n;
}

ORC_API int64_t orc_count_nnz(int64_t count, const float *A)
{
    int64_t n = 0;
    for (int64_t k = 0; k < count; k++) n += (A[k] != 0.0f);
    return n;
}

/* ------------------------------------------------------------------------------------------
 * Host packers
 * ---------------------------------------------------------------------------------------- */

/* CSRMatrix::CSRMatrix  matrix_csr.cpp:5-23.  CSR of A^T: one list per output column i;
 * row_pointers has N entries and NO trailing sentinel (matrix_csr.cpp:10-11). */
ORC_API int orc_pack_csr(int M, int N, const float *A, int32_t *row_ptr /*N*/,
                         int32_t *col_idx, float *vals)
{
    int cur = 0;
    for (int i = 0; i < N; i++) {
        row_ptr[i] = cur;
        for (int j = 0; j < M; j++) {
            float v = A[(size_t)j * N + i];
            if (v != 0.0f) { vals[cur] = v; col_idx[cur] = j; cur++; }
        }
    }
    return cur;
}

/* TCSRMatrix::TCSRMatrix  tcsr.cpp:5-38.  32x32 tiles, slab-major (block_x outer, block_y
 * inner); per tile 32 words, word i = column block_x+i, bit j = row block_y+j; values in
 * the same order; blk_idx = exclusive prefix of tile nnz WITH sentinel. */
ORC_API int orc_pack_tcsr(int M, int N, const float *A, int32_t *blk_idx, uint32_t *bitmaps,
                          float *vals)
{
    size_t bit = 0;
    int vi = 0, t = 0;
    memset(bitmaps, 0, (size_t)M * N / 32 * sizeof(uint32_t));
    blk_idx[t++] = 0;
    for (int bx = 0; bx < N; bx += 32)
        for (int by = 0; by < M; by += 32) {
            for (int i = 0; i < 32; i++)
                for (int j = 0; j < 32; j++) {
                    float v = A[(size_t)(by + j) * N + (bx + i)];
                    if (v != 0.0f) {
                        vals[vi++] = v;
                        bitmaps[bit / 32] |= 1u << (bit % 32);
                    }
                    bit++;
                }
            blk_idx[t++] = vi;
        }
    return vi;
}

/* WSPMatrix::WSPMatrix  wsp.cpp:3-40.  Bit i*M+j <-> A[j][i]; values per column padded to
 * nz_max_m = max column nnz.  Call with vals == NULL to obtain nz_max_m first. */
ORC_API int orc_pack_wsp(int M, int N, const float *A, uint32_t *bitmaps, float *vals,
                         int nz_max_m_in)
{
    int nz_max = 0;
    if (bitmaps) memset(bitmaps, 0, (size_t)M * N / 32 * sizeof(uint32_t));
    if (vals) memset(vals, 0, (size_t)N * nz_max_m_in * sizeof(float));
    size_t bit = 0;
    for (int i = 0; i < N; i++) {
        int k = 0;
        for (int j = 0; j < M; j++) {
            float v = A[(size_t)j * N + i];
            if (v != 0.0f) {
                if (vals) vals[(size_t)i * nz_max_m_in + k] = v;
                if (bitmaps) bitmaps[bit / 32] |= 1u << (bit % 32);
                k++;
            }
            bit++;
        }
        if (k > nz_max) nz_max = k;
    }
    return nz_max;
}

/* ASPMatrix::ASPMatrix  asp.cpp:3-14.  Dense re-tiling: slab-major, each 32x32 tile row-major. */
ORC_API void orc_pack_asp(int M, int N, const float *A, float *vals)
{
    size_t k = 0;
    for (int bn = 0; bn < N; bn += 32)
        for (int bm = 0; bm < M; bm += 32)
            for (int i = 0; i < 32; i++)
                for (int j = 0; j < 32; j++)
                    vals[k++] = A[(size_t)(bm + i) * N + bn + j];
}

/* AWSPMatrix::AWSPMatrix  awsp.cpp:3-49.  Slab-major 32x32 tiles; word tile*32+i = ROW bm+i,
 * bit j = column bn+j; per-tile values row-major, every tile padded to nz_bk_max.
 * Call with vals == NULL to obtain nz_bk_max first. */
ORC_API int orc_pack_awsp(int M, int N, const float *A, uint32_t *bitmaps, float *vals,
                          int nz_bk_max_in)
{
    int bk_max = 0;
    size_t ntiles = (size_t)M * N / 1024;
    if (bitmaps) memset(bitmaps, 0, (size_t)M * N / 32 * sizeof(uint32_t));
    if (vals) memset(vals, 0, ntiles * nz_bk_max_in * sizeof(float));
    size_t bit = 0, tile = 0;
    for (int bn = 0; bn < N; bn += 32)
        for (int bm = 0; bm < M; bm += 32) {
            int k = 0;
            for (int i = 0; i < 32; i++)
                for (int j = 0; j < 32; j++) {
                    float v = A[(size_t)(bm + i) * N + (bn + j)];
                    if (v != 0.0f) {
                        if (vals) vals[tile * nz_bk_max_in + k] = v;
                        if (bitmaps) bitmaps[bit / 32] |= 1u << (bit % 32);
                        k++;
                    }
                    bit++;
                }
            if (k > bk_max) bk_max = k;
            tile++;
        }
    return bk_max;
}

/* AWSPRefMatrix::AWSPRefMatrix  awsp_ref.cpp:4-58.  Same bitmap as AWSP (word slab*M+row);
 * values per (slab, quarter-of-M) stream, each quarter padded to its max over slabs;
 * warp_nz_offset = inclusive prefix of the 4 maxima.  Call with vals == NULL to obtain
 * warp_nz_offset first. */
ORC_API void orc_pack_awsp_ref(int M, int N, const float *A, uint32_t *bitmaps, float *vals,
                               int32_t *warp_nz_offset /*4, in/out*/)
{
    int maxq[4] = {0, 0, 0, 0};
    int off[4] = {0, 0, 0, 0};
    int stride = 0;
    if (vals) {
        for (int w = 0; w < 4; w++) off[w] = warp_nz_offset[w];
        stride = off[3];
        memset(vals, 0, (size_t)(N / 32) * stride * sizeof(float));
    }
    if (bitmaps) memset(bitmaps, 0, (size_t)M * N / 32 * sizeof(uint32_t));
    size_t bit = 0;
    int slab = 0;
    for (int bn = 0; bn < N; bn += 32, slab++)
        for (int w = 0; w < 4; w++) {
            int k = 0;
            int base = (w == 0) ? 0 : off[w - 1];
            for (int r = 0; r < M / 4; r++)
                for (int i = 0; i < 32; i++) {
                    float v = A[(size_t)(M * w / 4 + r) * N + (bn + i)];
                    if (v != 0.0f) {
                        if (vals) vals[(size_t)slab * stride + base + k] = v;
                        if (bitmaps) bitmaps[bit / 32] |= 1u << (bit % 32);
                        k++;
                    }
                    bit++;
                }
            if (k > maxq[w]) maxq[w] = k;
        }
    if (!vals) {
        int s = 0;
        for (int w = 0; w < 4; w++) { s += maxq[w]; warp_nz_offset[w] = s; }
    }
}

/* ------------------------------------------------------------------------------------------
 * Kernel-side decodes, restated on the CPU.  `gpu_order` = 0 accumulates each output with a
 * plain multiply-add in ascending row order (so the result must equal orc_sgemv_dense bit for
 * bit on finite inputs: skipped terms are exact zeros); `gpu_order` = 1 reproduces the
 * reference kernel's own association (lane partials, shuffle tree, 4-warp smem sum) with
 * fmaf(), i.e. the bits the reference kernel produces when built with nvcc's default -fmad.
 * ---------------------------------------------------------------------------------------- */
static inline float mad(float a, float b, float c, int fused)
{
    return fused ? fmaf(a, b, c) : (a * b + c);
}

static inline int popc32(uint32_t v) { return __builtin_popcount(v); }

/* csr_naive_kernel  csr_naive.cu:13-22: thread per output, k sequential,
 * end = (idx == N-1) ? nnz : row_ptr[idx+1]. */
ORC_API void orc_csr_naive_gemv(int M, int N, const float *vals, int nnz, const int32_t *col_idx,
                                const int32_t *row_ptr, const float *x, float *y, int gpu_order)
{
    (void)M;
    for (int i = 0; i < N; i++) {
        int b = row_ptr[i], e = (i == N - 1) ? nnz : row_ptr[i + 1];
        float acc = 0.0f;
        for (int k = b; k < e; k++) acc = mad(x[col_idx[k]], vals[k], acc, gpu_order);
        y[i] = acc;
    }
}

/* csr_tiling_kernel  csr_tiling.cu:47-113: per slab (32 outputs), per 32-row step: tile
 * decompressed into a zero-filled 32x32 smem tile (74-89: value index = blk_idx[tile] +
 * exclusive-prefix(popc(words)) + popc(word & lanemask_lt)), then warp 0 accumulates all
 * 32 rows in ascending order including the zero entries (95-103). */
ORC_API void orc_csr_tiling_gemv(int M, int N, const float *vals, const int32_t *blk_idx,
                                 const uint32_t *bitmaps, const float *x, float *y, int gpu_order)
{
    for (int s = 0; s < N / 32; s++)
        for (int lane = 0; lane < 32; lane++) {
            float acc = 0.0f;
            for (int ks = 0; ks < M; ks += 32) {
                const uint32_t *bmp = bitmaps + (size_t)s * M + ks; /* csr_tiling.cu:57 */
                int begin = blk_idx[s * (M / 32) + ks / 32];        /* csr_tiling.cu:77 */
                int pre = 0;
                for (int t = 0; t < lane; t++) pre += popc32(bmp[t]);
                uint32_t w = bmp[lane]; /* word `lane` = this output column's 32 rows */
                for (int i = 0; i < 32; i++) {
                    float a = 0.0f;
                    if (w & (1u << i)) a = vals[begin + pre + popc32(w & ((1u << i) - 1u))];
                    acc = mad(x[ks + i], a, acc, gpu_order);
                }
            }
            y[s * 32 + lane] = acc;
        }
}

/* wsp_kernel_v0 / v1  wsp.cu:23-55 / 83-137: warp per output column; per 1024 rows lane l
 * holds bitmap word l; round i: lanes whose bit is set add x[out_bk+32i+lane] * value at
 * nz_max_m*col + cnt + popc(word_i & lanemask_lt); then xor-shuffle tree 16,8,4,2,1.
 * (Needs M % 1024 == 0 like the kernel.)  v1 accumulates in the same order. */
ORC_API void orc_wsp_gemv(int M, int N, int nz_max_m, const uint32_t *bitmaps, const float *vals,
                          const float *x, float *y, int gpu_order)
{
    for (int col = 0; col < N; col++) {
        float lane_sum[32];
        for (int l = 0; l < 32; l++) lane_sum[l] = 0.0f;
        int cnt = 0;
        float seq = 0.0f;
        for (int w = 0; w < M / 32; w++) {
            uint32_t word = bitmaps[(size_t)col * (M / 32) + w];
            for (int l = 0; l < 32; l++)
                if (word & (1u << l)) {
                    float a = vals[(size_t)nz_max_m * col + cnt + popc32(word & ((1u << l) - 1u))];
                    float xv = x[w * 32 + l];
                    if (gpu_order) lane_sum[l] = fmaf(xv, a, lane_sum[l]);
                    else seq = xv * a + seq;
                }
            cnt += popc32(word);
        }
        if (gpu_order) {
            for (int s = 16; s >= 1; s >>= 1) { /* wsp.cu:50-52: every lane adds its xor partner */
                float nxt[32];
                for (int l = 0; l < 32; l++) nxt[l] = lane_sum[l] + lane_sum[l ^ s];
                memcpy(lane_sum, nxt, sizeof nxt);
            }
            y[col] = lane_sum[0];
        } else {
            y[col] = seq;
        }
    }
}

/* 4-warp fixed-order reduction shared by asp/awsp kernels (e.g. asp.cu:30-40):
 * warp0 + w1 + w2 + w3. */
static inline float reduce4(const float q[4]) { return ((q[0] + q[1]) + q[2]) + q[3]; }

/* asp_kernel_v0 (asp.cu:13-27) and asp_kernel_v2 (asp.cu:123-198) on the ASPMatrix layout:
 * block = slab of 32 outputs, warp w owns rows [wM/4,(w+1)M/4), lane = column; rows with
 * x == 0 are skipped.  version 2 keeps two 32-row tiles in flight, so within each 64-row
 * step the order is (tile0 row i, tile1 row i) for i = 0..31 (asp.cu:153-159). */
ORC_API void orc_asp_gemv(int M, int N, const float *tiled, const float *x, float *y, int version,
                          int gpu_order)
{
    for (int s = 0; s < N / 32; s++)
        for (int lane = 0; lane < 32; lane++) {
            float q[4];
            float seq = 0.0f;
            for (int w = 0; w < 4; w++) {
                float sum = 0.0f;
                const float *Ap = tiled + (size_t)s * M * 32 + (size_t)(32 * (M / 4)) * w + lane;
                const float *Xp = x + (M / 4) * w;
                int rows = M / 4;
                if (gpu_order && version == 2) {
                    for (int bk = 0; bk < rows; bk += 64)
                        for (int i = 0; i < 32; i++)
                            for (int t = 0; t < 2; t++) {
                                int r = bk + t * 32 + i;
                                if (Xp[r] != 0.0f) sum = fmaf(Ap[(size_t)r * 32], Xp[r], sum);
                            }
                } else {
                    for (int r = 0; r < rows; r++)
                        if (Xp[r] != 0.0f) {
                            if (gpu_order) sum = fmaf(Ap[(size_t)r * 32], Xp[r], sum);
                            else seq = Xp[r] * Ap[(size_t)r * 32] + seq;
                        }
                }
                q[w] = sum;
            }
            y[s * 32 + lane] = gpu_order ? reduce4(q) : seq;
        }
}

/* awsp_kernel_v0 (awsp.cu:20-47; x test commented out at 35,41) and awsp_kernel_v2
 * (awsp.cu:199-303; loads predicated on x != 0 && bit) on the AWSPMatrix layout: value of
 * (tile, row i, col lane) at tile*nz_bk_max + sum_{r<i} popc(word_r) + popc(word_i & lt). */
ORC_API void orc_awsp_gemv(int M, int N, int nz_bk_max, const uint32_t *bitmaps, const float *vals,
                           const float *x, float *y, int version, int gpu_order)
{
    for (int s = 0; s < N / 32; s++)
        for (int lane = 0; lane < 32; lane++) {
            float q[4];
            float seq = 0.0f;
            uint32_t lt = (1u << lane) - 1u, cur = 1u << lane;
            for (int w = 0; w < 4; w++) {
                float sum = 0.0f;
                int tile0 = s * (M / 32) + (M / 32) * w / 4; /* awsp.cu:22 */
                int ntile = M / 128;
                /* per-tile contribution of row i, as the kernel would add it */
                #define AWSP_TERM(T, I, OUT, HAVE)                                             \
                    do {                                                                       \
                        const uint32_t *bw = bitmaps + (size_t)(T) * 32;                       \
                        int pre = 0;                                                           \
                        for (int r_ = 0; r_ < (I); r_++) pre += popc32(bw[r_]);                \
                        uint32_t wd = bw[(I)];                                                 \
                        HAVE = (wd & cur) != 0;                                                \
                        OUT = HAVE ? vals[(size_t)(T) * nz_bk_max + pre + popc32(wd & lt)] : 0.0f; \
                    } while (0)
                if (gpu_order && version == 2) {
                    /* two tiles in flight; rows are still consumed in ascending order
                     * (awsp.cu:243-262: `for idx` outer, `for i` inner).  The main loop adds
                     * only when x_calc != 0 (251-252); the two tail stages, i.e. the last two
                     * 64-row steps, add unconditionally (279, 301). */
                    for (int t = 0; t < ntile; t += 2)
                        for (int idx = 0; idx < 2; idx++)
                            for (int i = 0; i < 32; i++) {
                                float a; int have;
                                AWSP_TERM(tile0 + t + idx, i, a, have);
                                float xv = x[(M / 4) * w + (t + idx) * 32 + i];
                                if (xv == 0.0f) a = 0.0f; /* is_load false -> A_buf = 0 (258-259) */
                                int tail = (t + 2 >= ntile - 2);
                                if (tail || xv != 0.0f) sum = fmaf(a, xv, sum);
                                (void)have;
                            }
                } else {
                    for (int t = 0; t < ntile; t++)
                        for (int i = 0; i < 32; i++) {
                            float a; int have;
                            AWSP_TERM(tile0 + t, i, a, have);
                            float xv = x[(M / 4) * w + t * 32 + i];
                            if (!have) continue;
                            if (gpu_order) sum = fmaf(xv, a, sum); /* v0: no x test (35,41) */
                            else if (xv != 0.0f) seq = xv * a + seq;
                        }
                }
                #undef AWSP_TERM
                q[w] = sum;
            }
            y[s * 32 + lane] = gpu_order ? reduce4(q) : seq;
        }
}

/* awsp_ref_kernel  awsp_ref.cu:20-184 on the AWSPRefMatrix layout: stream base
 * warp_offset[3]*slab + (w ? warp_offset[w-1] : 0) (20-25), running A_ptr += popc(word) per
 * row (54), bitmap word slab*M + row.  Pipeline order = awsp v2 (two tiles in flight). */
ORC_API void orc_awsp_ref_gemv(int M, int N, const uint32_t *bitmaps, const float *vals,
                               const int32_t *warp_offset, const float *x, float *y,
                               int gpu_order)
{
    int stride = warp_offset[3];
    for (int s = 0; s < N / 32; s++) {
        /* per (slab, quarter): exclusive prefix of popc over the quarter's rows */
        for (int lane = 0; lane < 32; lane++) {
            float q[4];
            float seq = 0.0f;
            uint32_t lt = (1u << lane) - 1u, cur = 1u << lane;
            for (int w = 0; w < 4; w++) {
                float sum = 0.0f;
                const float *Ap = vals + (size_t)stride * s + (w == 0 ? 0 : warp_offset[w - 1]);
                const uint32_t *Bp = bitmaps + (size_t)s * M + (size_t)(M / 4) * w;
                int rows = M / 4;
                if (gpu_order) {
                    /* need prefix at arbitrary row: precompute */
                    int *pre = (int *)malloc(sizeof(int) * (size_t)(rows + 1));
                    pre[0] = 0;
                    for (int r = 0; r < rows; r++) pre[r + 1] = pre[r] + popc32(Bp[r]);
                    for (int bk = 0; bk < rows; bk += 64)
                        for (int idx = 0; idx < 2; idx++)
                            for (int i = 0; i < 32; i++) {
                                int r = bk + idx * 32 + i;
                                float xv = x[(M / 4) * w + r];
                                uint32_t wd = Bp[r];
                                float a = ((wd & cur) && xv != 0.0f) ? Ap[pre[r] + popc32(wd & lt)] : 0.0f;
                                sum = fmaf(a, xv, sum); /* unconditional: awsp_ref.cu:88,103,136,151,163,169 */
                            }
                    free(pre);
                } else {
                    int p = 0;
                    for (int r = 0; r < rows; r++) {
                        uint32_t wd = Bp[r];
                        float xv = x[(M / 4) * w + r];
                        if ((wd & cur) && xv != 0.0f) seq = xv * Ap[p + popc32(wd & lt)] + seq;
                        p += popc32(wd);
                    }
                }
                q[w] = sum;
            }
            y[s * 32 + lane] = gpu_order ? reduce4(q) : seq;
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * CSR(A^T) product for shapes with no dense form (BASELINE configs 4/5): csr_naive.cu:14-22
 * semantics with 64-bit pointers and an N+1 sentinel; optional OpenMP for the CPU baseline.
 * ---------------------------------------------------------------------------------------- */
ORC_API void orc_csc_gemv(int64_t N, const int64_t *col_ptr, const int32_t *row_idx,
                          const float *vals, const float *x, float *y, int threads)
{
    (void)threads;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 256) num_threads(threads > 0 ? threads : 1)
#endif
    for (int64_t i = 0; i < N; i++) {
        float acc = 0.0f;
        for (int64_t k = col_ptr[i]; k < col_ptr[i + 1]; k++) acc += x[row_idx[k]] * vals[k];
        y[i] = acc;
    }
}

/* Dense -> CSR(A^T) with sentinel, 64-bit pointers (same traversal as matrix_csr.cpp:10-22). */
ORC_API int64_t orc_dense_to_csc(int M, int N, const float *A, int64_t *col_ptr, int32_t *row_idx,
                                 float *vals)
{
    int64_t cur = 0;
    for (int i = 0; i < N; i++) {
        col_ptr[i] = cur;
        for (int j = 0; j < M; j++) {
            float v = A[(size_t)j * N + i];
            if (v != 0.0f) {
                if (vals) { vals[cur] = v; row_idx[cur] = j; }
                cur++;
            }
        }
    }
    col_ptr[N] = cur;
    return cur;
}
