// ref_shim_cpu.cpp — TEST INFRASTRUCTURE.  C-ABI window onto the UNMODIFIED reference host
// code, compiled from where it lies under $(REF) (= /root/reference) by oracle/Makefile into
// oracle/_ref/libspmv_ref_cpu.so.  Nothing of the reference is copied into this repository:
// this file only #includes / links the reference's own translation units.
//
// Used to (1) pin oracle/spmv_oracle.c and (2) serve as the `--impl reference` CPU arm of
// bench.py (kind "reference": the reference's own SgemvCPU, tester.cpp:36-45).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <iomanip>
#include <iostream>
#include <random>
#include <string>
#include <vector>
#include <assert.h>

// The reference class keeps its inputs and SgemvCPU private (tester.hpp:31-57).  All standard
// headers tester.hpp/tester.cpp need are already included above (their include guards make the
// re-includes no-ops), so redefining the access keyword only affects the reference class.
#define private public
#include "tester.hpp"
#undef private

#include "matrix_csr.hpp"
#include "tcsr.hpp"
#include "wsp.hpp"
#include "asp.hpp"
#include "awsp.hpp"
#include "awsp_ref.hpp"

// tester.cpp's SgemvGPU (tester.cpp:54-63) references the GPU launchers; the CPU shim never
// calls SgemvGPU, so satisfy the linker with traps.
#define TRAP(name, ...) void name(__VA_ARGS__) { fprintf(stderr, #name ": GPU launcher not in the CPU shim\n"); abort(); }
TRAP(cublas_gemv_gpu, int, int, float *, float *, float *)
TRAP(wsp_gemv_gpu, int, int, float *, float *, float *, int)
TRAP(asp_gemv_gpu, int, int, float *, float *, float *, int)
TRAP(awsp_gemv_gpu, int, int, float *, float *, float *, int)
TRAP(awsp_ref_gemv_gpu, int, int, float *, float *, float *)

template <class T> static T *dup(const T *src, size_t n)
{
    T *p = (T *)malloc(n * sizeof(T) + 1);
    if (n) memcpy(p, src, n * sizeof(T));
    return p;
}

extern "C" {

// SparseSgemvTester::SgemvCPU, tester.cpp:36-45, run on caller-provided inputs.
void ref_sgemv_cpu(int M, int N, const float *A, const float *x, float *y)
{
    SparseSgemvTester t(M, N);
    t.A_host = const_cast<float *>(A);
    t.X_host = const_cast<float *>(x);
    t.SgemvCPU();
    memcpy(y, t.Y_cpu_host, sizeof(float) * (size_t)N);
    t.A_host = nullptr; // not ours to free (tester.hpp:15-28 frees non-null members)
    t.X_host = nullptr;
}

struct ref_packed {
    int32_t *i32_a; int64_t n_i32_a;
    int32_t *i32_b; int64_t n_i32_b;
    uint32_t *u32;  int64_t n_u32;
    float *f32;     int64_t n_f32;
    int32_t aux[4];
};

// layout ids match spmv_layout_t in include/spmv_b200.h
int ref_pack(int layout, int M, int N, const float *A_in, ref_packed *o)
{
    float *A = const_cast<float *>(A_in);
    memset(o, 0, sizeof *o);
    switch (layout) {
    case 0: { CSRMatrix m(M, N, A);
        o->i32_a = dup(m.GetRowPtrs(), m.RowPtrsSize()); o->n_i32_a = m.RowPtrsSize();
        o->i32_b = dup(m.GetColIdxs(), m.ColIdxsSize()); o->n_i32_b = m.ColIdxsSize();
        o->f32 = dup(m.GetValues(), m.ValuesSize());     o->n_f32 = m.ValuesSize();
        return 0; }
    case 1: { TCSRMatrix m(M, N, A);
        o->i32_a = dup(m.GetBlkIdx(), m.BlkIdxSize());   o->n_i32_a = m.BlkIdxSize();
        o->u32 = dup(m.GetBitmaps(), m.BitmapsSize());   o->n_u32 = m.BitmapsSize();
        o->f32 = dup(m.GetValues(), m.ValuesSize());     o->n_f32 = m.ValuesSize();
        return 0; }
    case 2: { WSPMatrix m(M, N, A);
        o->u32 = dup(m.GetBitmaps(), m.BitmapsSize());   o->n_u32 = m.BitmapsSize();
        o->f32 = dup(m.GetValues(), m.ValuesSize());     o->n_f32 = m.ValuesSize();
        o->aux[0] = m.nz_max_m; o->aux[1] = m.nz_max_n;
        return 0; }
    case 3: { ASPMatrix m(M, N, A);
        o->f32 = dup(m.GetValues(), m.ValuesSize());     o->n_f32 = m.ValuesSize();
        return 0; }
    case 4: { AWSPMatrix m(M, N, A);
        o->u32 = dup(m.GetBitmaps(), m.BitmapsSize());   o->n_u32 = m.BitmapsSize();
        o->f32 = dup(m.GetValues(), m.ValuesSize());     o->n_f32 = m.ValuesSize();
        o->aux[0] = m.nz_bk_max_;
        return 0; }
    case 5: { AWSPRefMatrix m(M, N, A);
        o->i32_a = dup(m.GetWarpNZOffset(), 4);          o->n_i32_a = 4;
        o->u32 = dup(m.GetBitmaps(), m.BitmapsSize());   o->n_u32 = m.BitmapsSize();
        o->f32 = dup(m.GetValues(), m.ValuesSize());     o->n_f32 = m.ValuesSize();
        return 0; }
    }
    return -1;
}

void ref_packed_free(ref_packed *o)
{
    free(o->i32_a); free(o->i32_b); free(o->u32); free(o->f32);
    memset(o, 0, sizeof *o);
}

} // extern "C"
