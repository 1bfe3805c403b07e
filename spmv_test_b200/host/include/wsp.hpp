// wsp.hpp — drop-in for the reference's WSPMatrix (src/include/wsp.hpp:4-20): per-column
// bitmaps + ELL-padded values, public nz_max_m / nz_max_n (wsp.cpp:3-40).
#pragma once
#include <cstdint>
#include <iostream>
#include <vector>

#include "ref_layout.hpp"

class WSPMatrix {
public:
    WSPMatrix(int M, int N, float *matrix)
    {
        data_.Build(2, M, N, matrix);
        nz_max_m = data_.aux[0];
        nz_max_n = data_.aux[1];
    }

    auto GetBitmaps() -> uint32_t * { return data_.words.data(); }
    auto GetValues() -> float * { return data_.reals.data(); }
    auto BitmapsSize() -> int { return static_cast<int>(data_.words.size()); }
    auto ValuesSize() -> int { return static_cast<int>(data_.reals.size()); }

    int nz_max_m, nz_max_n;

private:
    spmv_host::PackedLayout data_;
};
