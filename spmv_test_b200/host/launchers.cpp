// launchers.cpp — the reference's ten launchers (src/include/kernel.hpp:8-17) over the C-ABI.
//
// A reference launcher is: pack A on the host, cudaMalloc, H2D, one timed launch, D2H,
// cudaFree (e.g. awsp.cu:319-388).  Same life cycle here, one plan per call; callers that
// want pack-once / run-many use the plan API of include/spmv_b200.h directly.
#include <chrono>
#include <cstdint>
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

// cublas_gemv_gpu (the reference's dense comparator, cublas.cu:4-44) lives in cublas_comparator.cpp:
// it calls cublasSgemv like the reference does and is linked into the harness, not into the library.

// Multi-GPU form of the awsp launcher (the reference is single-GPU): A's columns are
// cut into one slab per visible device, every device packs and keeps its slab, one call runs all of
// them (spmv_mg_group_run_host: peer stores into every device's y, in-kernel arrival) and Y comes
// back whole.  With one device it is the plain launcher.
void awsp_mg_gemv_gpu(int M, int N, float *A_host, float *X_host, float *Y_host)
{
    int n_dev = spmv_device_count();
    if (n_dev > 8) n_dev = 8;
    while (n_dev > 1 && (N / 32) < n_dev) n_dev--;        // at least one 32-column unit per device
    if (n_dev <= 1) { run_once("awsp_kernel (1 device)", SPMV_AWSP, M, N, A_host, X_host, Y_host); return; }
    int devices[8];
    int64_t bounds[9];
    for (int d = 0; d < n_dev; d++) devices[d] = d;
    if (spmv_partition_columns(N, n_dev, 32, nullptr, bounds) != SPMV_OK) die("awsp_mg partition");
    spmv_mg_t *groups[8] = {};
    spmv_plan_t *plans[8] = {};
    if (spmv_mg_create_group(M, N, n_dev, devices, groups) != SPMV_OK) die("awsp_mg group");
    for (int d = 0; d < n_dev; d++) {
        if (spmv_set_device(devices[d]) != SPMV_OK) die("awsp_mg device");
        const int64_t w = bounds[d + 1] - bounds[d];
        if (spmv_plan_create_dense(SPMV_AWSP, M, w, A_host + bounds[d], N, nullptr, &plans[d]) != SPMV_OK) die("awsp_mg plan");
        if (spmv_mg_add_plan(groups[d], plans[d], bounds[d]) != SPMV_OK) die("awsp_mg add");
    }
    if (spmv_mg_group_run_host(groups, n_dev, X_host, Y_host) != SPMV_OK) die("awsp_mg run");   // warm-up: module load on every device
    const auto t0 = std::chrono::steady_clock::now();
    if (spmv_mg_group_run_host(groups, n_dev, X_host, Y_host) != SPMV_OK) die("awsp_mg run");
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    std::cout << "awsp_kernel x " << n_dev << " devices (host x -> all devices, kernels, arrival, y slices -> host) took " << ms
              << " ms" << std::endl;
    for (int d = 0; d < n_dev; d++) {
        spmv_set_device(devices[d]);
        spmv_mg_destroy(groups[d]);
        spmv_plan_destroy(plans[d]);
    }
    spmv_set_device(devices[0]);
}
