// kernel.hpp — drop-in for the reference's launcher header (src/include/kernel.hpp:8-17).
//
// Same ten free functions, same argument meaning: host pointers in, host pointer out,
// synchronous, A_host dense row-major M x N, Y_host fully overwritten (N floats); `version`
// selects a kernel generation in the reference (wsp 0/1, asp and awsp 0/1/2) and an unknown
// version launches nothing (wsp.cu:187-188) — kept.  Every launcher packs A into this
// library's device format, uploads, runs the sm_100a kernel once, downloads and frees, and
// prints "<what> took <ms> ms" to stdout like the reference's TIME_KERNEL (kernel.hpp:31-48).
// On a CUDA failure it prints "CUDA error ..." to stderr and exits, like CUDA_CHECK (21-28).
//
// Mapping to the C-ABI variants (include/spmv_b200.h):
//   wsp, csr_naive                 -> SPMV_WSP       asp, naive, tiling, cublas -> SPMV_ASP (dense A)
//   awsp, awsp_ref, wsp_sm         -> SPMV_AWSP      csr_tiling                 -> SPMV_TCSR
// cublas_gemv_gpu is the reference's dense comparator: cublasSgemv (host/cublas_comparator.cpp), linked into
// the harness only.  awsp_mg_gemv_gpu is an addition: the awsp launcher over every visible device.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <iostream>

void tiling_gemv_gpu(int M, int N, float *A_host, float *X_host, float *Y_host);
void cublas_gemv_gpu(int M, int N, float *A, float *X, float *Y);
void naive_gemv_gpu(int M, int N, float *A_host, float *X_host, float *Y_host);
void csr_naive_gemv_gpu(int M, int N, float *A_host, float *X_host, float *Y_host);
void csr_tiling_gemv_gpu(int M, int N, float *A_host, float *X_host, float *Y_host);
void wsp_gemv_gpu(int M, int N, float *A_host, float *X_host, float *Y_host, int version);
void asp_gemv_gpu(int M, int N, float *A_host, float *X_host, float *Y_host, int version);
void awsp_gemv_gpu(int M, int N, float *A_host, float *X_host, float *Y_host, int version);
void awsp_ref_gemv_gpu(int M, int N, float *A_host, float *X_host, float *Y_host);
void wsp_sm_gemv_gpu(int M, int N, float *A_host, float *X_host, float *Y_host);
void awsp_mg_gemv_gpu(int M, int N, float *A_host, float *X_host, float *Y_host);   // not in the reference: all visible GPUs

// The reference's helper macros, for translation units that include the CUDA runtime
// themselves; this header does not need it.
#if defined(__CUDACC__) || defined(CUDART_VERSION)
#define CUDA_CHECK(call)                                                                            \
    do {                                                                                            \
        const cudaError_t spmv_err_ = (call);                                                       \
        if (spmv_err_ != cudaSuccess) {                                                             \
            fprintf(stderr, "CUDA error %s:%d: %s\n", __FILE__, __LINE__, cudaGetErrorString(spmv_err_)); \
            exit(EXIT_FAILURE);                                                                     \
        }                                                                                           \
    } while (0)

#define TIME_KERNEL(kernel_call)                                                                    \
    do {                                                                                            \
        cudaEvent_t spmv_t0_, spmv_t1_;                                                             \
        CUDA_CHECK(cudaEventCreate(&spmv_t0_));                                                     \
        CUDA_CHECK(cudaEventCreate(&spmv_t1_));                                                     \
        CUDA_CHECK(cudaEventRecord(spmv_t0_));                                                      \
        (kernel_call);                                                                              \
        CUDA_CHECK(cudaEventRecord(spmv_t1_));                                                      \
        CUDA_CHECK(cudaEventSynchronize(spmv_t1_));                                                 \
        float spmv_ms_ = 0.0f;                                                                      \
        CUDA_CHECK(cudaEventElapsedTime(&spmv_ms_, spmv_t0_, spmv_t1_));                            \
        std::cout << #kernel_call << " took " << spmv_ms_ << " ms" << std::endl;                    \
        cudaEventDestroy(spmv_t0_);                                                                 \
        cudaEventDestroy(spmv_t1_);                                                                 \
    } while (0)
#endif
