// tester.hpp — drop-in for the reference's harness class (src/include/tester.hpp:10-58).
// Same public surface: SparseSgemvTester(int m, int n), RunTest().  Inputs follow the
// reference's generator (Bernoulli(0.5) mask x U(-1,1), tester.cpp:103-121, 151-167); set
// SPMV_SEED to make them reproducible, SPMV_SPARSITY_A / SPMV_SPARSITY_X to change the masks,
// SPMV_STRICT=1 to turn a mismatch into a non-zero exit code (the reference only prints).
#pragma once

#include <cstdint>
#include <functional>
#include <string>
#include <vector>

class SparseSgemvTester {
public:
    SparseSgemvTester(int m, int n);
    ~SparseSgemvTester();
    SparseSgemvTester(const SparseSgemvTester &) = delete;
    SparseSgemvTester &operator=(const SparseSgemvTester &) = delete;

    auto RunTest() -> void;

private:
    using Launcher = std::function<void(int, int, float *, float *, float *)>;
    struct Entry { std::string name; Launcher run; };

    int m_, n_;
    float *A_host = nullptr;
    float *X_host = nullptr;
    float *Y_cpu_host = nullptr;
    std::vector<float *> Y_gpu_hosts;
    std::vector<Entry> registry_;
    long mismatches_ = 0;

    auto GetRandomMatrix() -> void;
    auto GetRandomVector() -> void;
    auto SgemvCPU() -> void;     // dense sequential fp32 reference (tester.cpp:36-45)
    auto SgemvGPU() -> void;     // every registered launcher, fresh y each (tester.cpp:47-72)
    auto CompareY() -> void;     // abs 1e-3 gate (tester.cpp:74-88)
};
