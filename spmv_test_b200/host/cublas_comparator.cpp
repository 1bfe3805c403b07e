// cublas_comparator.cpp — the reference's dense comparator (src/kernels/cublas.cu:4-44): y = x*A as
// cublasSgemv(CUBLAS_OP_N, N, M, 1, A, lda = N, x, 1, 0, y, 1) on the row-major A seen as a
// column-major N x M matrix.  A LIBRARY call kept only as an independent yardstick next to the
// sparse kernels: it is linked into the harness executable, never into
// libspmv_b200.so, and nothing in the product path calls it.
#include <cstdio>
#include <cstdlib>
#include <iostream>

#include <cublas_v2.h>
#include <cuda_runtime.h>

#include "kernel.hpp"

namespace {
void cublas_check(cublasStatus_t s, const char *what)
{
    if (s != CUBLAS_STATUS_SUCCESS) {
        fprintf(stderr, "cuBLAS error in %s: %d\n", what, (int)s);
        exit(EXIT_FAILURE);
    }
}
} // namespace

void cublas_gemv_gpu(int M, int N, float *A, float *X, float *Y)
{
    float *dA = nullptr, *dx = nullptr, *dy = nullptr;
    CUDA_CHECK(cudaMalloc(&dA, sizeof(float) * (size_t)M * N));
    CUDA_CHECK(cudaMalloc(&dx, sizeof(float) * (size_t)M));
    CUDA_CHECK(cudaMalloc(&dy, sizeof(float) * (size_t)N));
    CUDA_CHECK(cudaMemcpy(dA, A, sizeof(float) * (size_t)M * N, cudaMemcpyHostToDevice));
    CUDA_CHECK(cudaMemcpy(dx, X, sizeof(float) * (size_t)M, cudaMemcpyHostToDevice));
    cublasHandle_t h;
    cublas_check(cublasCreate(&h), "cublasCreate");
    const float one = 1.0f, zero = 0.0f;
    cublas_check(cublasSgemv(h, CUBLAS_OP_N, N, M, &one, dA, N, dx, 1, &zero, dy, 1), "cublasSgemv (warm-up)");
    CUDA_CHECK(cudaDeviceSynchronize());
    TIME_KERNEL(cublas_check(cublasSgemv(h, CUBLAS_OP_N, N, M, &one, dA, N, dx, 1, &zero, dy, 1), "cublasSgemv"));
    CUDA_CHECK(cudaMemcpy(Y, dy, sizeof(float) * (size_t)N, cudaMemcpyDeviceToHost));
    cublasDestroy(h);
    cudaFree(dA); cudaFree(dx); cudaFree(dy);
}
