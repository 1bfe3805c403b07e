// asp.hpp — drop-in for the reference's ASPMatrix (src/include/asp.hpp:3-11): dense A re-tiled
// into 32x32 tiles, slab-major (asp.cpp:3-14).
#pragma once
#include <vector>

#include "ref_layout.hpp"

class ASPMatrix {
public:
    ASPMatrix(int M, int N, float *matrix) { data_.Build(3, M, N, matrix); }

    auto GetValues() -> float * { return data_.reals.data(); }
    auto ValuesSize() -> int { return static_cast<int>(data_.reals.size()); }

private:
    spmv_host::PackedLayout data_;
};
