// tcsr.hpp — drop-in for the reference's TCSRMatrix (src/include/tcsr.hpp:4-23): 32x32 tiled
// bitmap-CSR, blk_idx with sentinel (tcsr.cpp:5-38).
#pragma once
#include <cstdint>
#include <iostream>
#include <vector>

#include "ref_layout.hpp"

class TCSRMatrix {
public:
    TCSRMatrix(int m, int n, float *matrix) { data_.Build(1, m, n, matrix); }

    auto GetBlkIdx() -> int * { return data_.ints_a.data(); }
    auto BlkIdxSize() -> int { return static_cast<int>(data_.ints_a.size()); }
    auto GetBitmaps() -> uint32_t * { return data_.words.data(); }
    auto BitmapsSize() -> int { return static_cast<int>(data_.words.size()); }
    auto GetValues() -> float * { return data_.reals.data(); }
    auto ValuesSize() -> int { return static_cast<int>(data_.reals.size()); }

private:
    spmv_host::PackedLayout data_;
};
