// launchers.cpp — the reference's ten launchers (src/include/kernel.hpp:8-17) over the C-ABI.
//
// A reference launcher is: pack A on the host, cudaMalloc, H2D, one timed launch, D2H,
// cudaFree (e.g. awsp.cu:319-388).  Same life cycle here, one plan per call; callers that
// want pack-once / run-many use the plan API of include/spmv_b200.h directly.
#include <cstdio>
#include <cstdlib>
#include <iostream>

#include "kernel.hpp"
#include "spmv_b200.h"

namespace {

[[noreturn]] void die(const char *where)
{
    // same shape as the reference's CUDA_CHECK message (kernel.hpp:24-26)
    fprintf(stderr, "CUDA error %s: %s\n", where, spmv_last_error());
    exit(EXIT_FAILURE);
}

void run_once(const char *label, int variant, int M, int N, float *A, float *X, float *Y)
{
    spmv_plan_t *plan = nullptr;
    if (spmv_plan_create_dense(variant, M, N, A, N, nullptr, &plan) != SPMV_OK) die(label);
    spmv_plan_info_t info;
    spmv_plan_info(plan, &info);
    float ms = 0.0f;
    if (spmv_run_host(plan, X, Y, &ms) != SPMV_OK) die(label);
    std::cout << label << "<<<(" << info.grid_x << "," << info.grid_y << "), " << info.block << ">>> took " << ms
              << " ms" << std::endl;
    spmv_plan_destroy(plan);
}

} // namespace

void wsp_gemv_gpu(int M, int N, float *A_host, float *X_host, float *Y_host, int version)
{
    if (version != 0 && version != 1) return;   // reference: unknown version launches nothing
    run_once(version ? "wsp_kernel_v1" : "wsp_kernel_v0", SPMV_WSP, M, N, A_host, X_host, Y_host);
}

void asp_gemv_gpu(int M, int N, float *A_host, float *X_host, float *Y_host, int version)
{
    static const char *names[3] = {"asp_kernel_v0", "asp_kernel_v1", "asp_kernel_v2"};
    if (version < 0 || version > 2) return;
    run_once(names[version], SPMV_ASP, M, N, A_host, X_host, Y_host);
}

void awsp_gemv_gpu(int M, int N, float *A_host, float *X_host, float *Y_host, int version)
{
    static const char *names[3] = {"awsp_kernel_v0", "awsp_kernel_v1", "awsp_kernel_v2"};
    if (version < 0 || version > 2) return;
    run_once(names[version], SPMV_AWSP, M, N, A_host, X_host, Y_host);
}

void awsp_ref_gemv_gpu(int M, int N, float *A_host, float *X_host, float *Y_host)
{
    run_once("awsp_ref_kernel", SPMV_AWSP, M, N, A_host, X_host, Y_host);
}

void wsp_sm_gemv_gpu(int M, int N, float *A_host, float *X_host, float *Y_host)
{
    run_once("wsp_sm_kernel", SPMV_AWSP, M, N, A_host, X_host, Y_host);
}

void csr_naive_gemv_gpu(int M, int N, float *A_host, float *X_host, float *Y_host)
{
    run_once("csr_naive_kernel", SPMV_WSP, M, N, A_host, X_host, Y_host);
}

void csr_tiling_gemv_gpu(int M, int N, float *A_host, float *X_host, float *Y_host)
{
    run_once("csr_tiling_kernel", SPMV_TCSR, M, N, A_host, X_host, Y_host);
}

// dense comparators: the dense-A kernel of the asp variant
void naive_gemv_gpu(int M, int N, float *A_host, float *X_host, float *Y_host)
{
    run_once("naive_kernel", SPMV_ASP, M, N, A_host, X_host, Y_host);
}

void tiling_gemv_gpu(int M, int N, float *A_host, float *X_host, float *Y_host)
{
    run_once("tiling_kernel", SPMV_ASP, M, N, A_host, X_host, Y_host);
}

void cublas_gemv_gpu(int M, int N, float *A, float *X, float *Y)
{
    run_once("dense_sgemv", SPMV_ASP, M, N, A, X, Y);
}
