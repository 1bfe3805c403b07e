// Harness entry point.  Without arguments it is the reference's self-check — one 4096 x 4096
// problem (reference test/main.cpp:4) — and `sparse_sgemv M N` runs the same check on another
// shape (both multiples of 32, as the tester asserts), e.g. `sparse_sgemv 4096 14336` for the
// decode up-projection.  SPMV_SEED / SPMV_SPARSITY_A / SPMV_SPARSITY_X / SPMV_STRICT: see tester.cpp.
#include <cstdio>
#include <cstdlib>

#include "tester.hpp"

int main(int argc, char **argv)
{
    int rows = 4096, cols = 4096;
    if (argc == 3) {
        rows = std::atoi(argv[1]);
        cols = std::atoi(argv[2]);
    } else if (argc != 1) {
        std::fprintf(stderr, "usage: %s [M N]\n", argv[0]);
        return 2;
    }
    if (rows <= 0 || cols <= 0 || rows % 32 || cols % 32) {
        std::fprintf(stderr, "M and N must be positive multiples of 32\n");
        return 2;
    }
    SparseSgemvTester harness(rows, cols);
    harness.RunTest();
    return 0;
}
