// matrix_csr.hpp — drop-in for the reference's CSRMatrix (src/include/matrix_csr.hpp:4-25).
// CSR of A^T: one (value, row) list per output column, row_pointers WITHOUT the N+1 sentinel
// (matrix_csr.cpp:8-22).  Same constructor and accessors; storage is a PackedLayout.
#pragma once
#include <iostream>
#include <vector>

#include "ref_layout.hpp"

class CSRMatrix {
public:
    CSRMatrix(int m, int n, float *matrix) : rows_(m), cols_(n) { data_.Build(0, m, n, matrix); }

    auto GetRowPtrs() -> int * { return data_.ints_a.data(); }
    auto RowPtrsSize() -> int { return static_cast<int>(data_.ints_a.size()); }
    auto GetColIdxs() -> int * { return data_.ints_b.data(); }
    auto ColIdxsSize() -> int { return static_cast<int>(data_.ints_b.size()); }
    auto GetValues() -> float * { return data_.reals.data(); }
    auto ValuesSize() -> int { return static_cast<int>(data_.reals.size()); }
    void PrintCSR();

private:
    int rows_, cols_;
    spmv_host::PackedLayout data_;
};
