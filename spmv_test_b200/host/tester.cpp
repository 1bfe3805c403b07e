// tester.cpp — the harness behind test/main.cpp (reference src/tester.cpp:6-221, rewritten).
#include "tester.hpp"

#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <random>

#include "kernel.hpp"

namespace {

double env_double(const char *name, double fallback)
{
    const char *s = std::getenv(name);
    return s ? std::atof(s) : fallback;
}

std::mt19937 make_rng(unsigned salt)
{
    // the reference seeds from std::random_device (tester.cpp:107,155); SPMV_SEED pins it
    if (const char *s = std::getenv("SPMV_SEED")) return std::mt19937((unsigned)std::strtoul(s, nullptr, 10) + salt);
    std::random_device rd;
    return std::mt19937(rd());
}

void fill_sparse(float *dst, size_t count, double sparsity, std::mt19937 &rng)
{
    std::uniform_real_distribution<float> coin(0.0f, 1.0f), value(-1.0f, 1.0f);
    for (size_t k = 0; k < count; k++) dst[k] = coin(rng) > sparsity ? value(rng) : 0.0f;
}

} // namespace

SparseSgemvTester::SparseSgemvTester(int m, int n) : m_(m), n_(n)
{
    assert(m % 32 == 0);     // tester.cpp:9-10
    assert(n % 32 == 0);
    // the reference's run list, in its order (tester.cpp:54-63) ...
    registry_ = {
        {"cublas", [](int M, int N, float *A, float *X, float *Y) { cublas_gemv_gpu(M, N, A, X, Y); }},
        {"wsp v0", [](int M, int N, float *A, float *X, float *Y) { wsp_gemv_gpu(M, N, A, X, Y, 0); }},
        {"wsp v1", [](int M, int N, float *A, float *X, float *Y) { wsp_gemv_gpu(M, N, A, X, Y, 1); }},
        {"asp v2", [](int M, int N, float *A, float *X, float *Y) { asp_gemv_gpu(M, N, A, X, Y, 2); }},
        {"awsp v0", [](int M, int N, float *A, float *X, float *Y) { awsp_gemv_gpu(M, N, A, X, Y, 0); }},
        {"awsp v1", [](int M, int N, float *A, float *X, float *Y) { awsp_gemv_gpu(M, N, A, X, Y, 1); }},
        {"awsp v2", [](int M, int N, float *A, float *X, float *Y) { awsp_gemv_gpu(M, N, A, X, Y, 2); }},
        {"awsp_ref", [](int M, int N, float *A, float *X, float *Y) { awsp_ref_gemv_gpu(M, N, A, X, Y); }},
        // ... plus the csr launchers it declares (kernel.hpp:11-12) but never runs
        {"csr_naive", [](int M, int N, float *A, float *X, float *Y) { csr_naive_gemv_gpu(M, N, A, X, Y); }},
        {"csr_tiling", [](int M, int N, float *A, float *X, float *Y) { csr_tiling_gemv_gpu(M, N, A, X, Y); }},
        // ... and the column-sharded form over every visible GPU (one device: the plain awsp launcher)
        {"awsp multi-gpu", [](int M, int N, float *A, float *X, float *Y) { awsp_mg_gemv_gpu(M, N, A, X, Y); }},
    };
}

SparseSgemvTester::~SparseSgemvTester()
{
    std::free(A_host);
    std::free(X_host);
    std::free(Y_cpu_host);
    for (float *y : Y_gpu_hosts) std::free(y);
}

auto SparseSgemvTester::RunTest() -> void
{
    GetRandomMatrix();
    GetRandomVector();
    SgemvCPU();
    SgemvGPU();
    CompareY();
    if (mismatches_ == 0) {
        std::cout << "========== OK ===========" << std::endl;
    } else {
        std::cout << "========== " << mismatches_ << " MISMATCHES ===========" << std::endl;
        if (env_double("SPMV_STRICT", 0.0) != 0.0) std::exit(EXIT_FAILURE);
    }
}

auto SparseSgemvTester::GetRandomMatrix() -> void
{
    A_host = static_cast<float *>(std::malloc(sizeof(float) * (size_t)m_ * n_));
    auto rng = make_rng(0);
    fill_sparse(A_host, (size_t)m_ * n_, env_double("SPMV_SPARSITY_A", 0.5), rng);
}

auto SparseSgemvTester::GetRandomVector() -> void
{
    X_host = static_cast<float *>(std::malloc(sizeof(float) * (size_t)m_));
    auto rng = make_rng(1);
    fill_sparse(X_host, (size_t)m_, env_double("SPMV_SPARSITY_X", 0.5), rng);
}

auto SparseSgemvTester::SgemvCPU() -> void
{
    Y_cpu_host = static_cast<float *>(std::malloc(sizeof(float) * (size_t)n_));
    // same arithmetic as the reference (one fp32 accumulator per output, rows ascending), but
    // swept row by row so the inner loop is unit stride; each y[i] sees the identical sequence
    // of additions, so the result is bit-identical to the column-by-column loop.
    for (int i = 0; i < n_; i++) Y_cpu_host[i] = 0.0f;
    for (int j = 0; j < m_; j++) {
        const float xj = X_host[j];
        const float *row = A_host + (size_t)j * n_;
        for (int i = 0; i < n_; i++) Y_cpu_host[i] += xj * row[i];
    }
}

auto SparseSgemvTester::SgemvGPU() -> void
{
    for (const Entry &e : registry_) {
        std::cout << "start to launch " << e.name << " kernel" << std::endl;
        float *y = static_cast<float *>(std::malloc(sizeof(float) * (size_t)n_));
        e.run(m_, n_, A_host, X_host, y);
        Y_gpu_hosts.push_back(y);
    }
}

auto SparseSgemvTester::CompareY() -> void
{
    const float tol = 1e-3f;   // tester.cpp:75
    for (size_t k = 0; k < Y_gpu_hosts.size(); k++)
        for (int i = 0; i < n_; i++) {
            const float d = std::fabs(Y_cpu_host[i] - Y_gpu_hosts[k][i]);
            if (!(d <= tol)) {
                if (mismatches_ < 64)
                    fprintf(stderr, "[GPU kernel %zu: %s] at [%d], cpu: %f, gpu: %f\n", k, registry_[k].name.c_str(), i,
                            Y_cpu_host[i], Y_gpu_hosts[k][i]);
                mismatches_++;
            }
        }
}
