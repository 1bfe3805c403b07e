// awsp_ref.hpp — drop-in for the reference's AWSPRefMatrix (src/include/awsp_ref.hpp:4-18):
// row bitmaps + per-(slab, quarter-of-M) value streams, warp_nz_offset (awsp_ref.cpp:4-58).
#pragma once
#include <cstdint>
#include <vector>

#include "ref_layout.hpp"

class AWSPRefMatrix {
public:
    AWSPRefMatrix(int M, int N, float *matrix) { data_.Build(5, M, N, matrix); }

    auto GetBitmaps() -> uint32_t * { return data_.words.data(); }
    auto GetValues() -> float * { return data_.reals.data(); }
    auto BitmapsSize() -> int { return static_cast<int>(data_.words.size()); }
    auto ValuesSize() -> int { return static_cast<int>(data_.reals.size()); }
    auto GetWarpNZOffset() -> int * { return data_.ints_a.data(); }

private:
    spmv_host::PackedLayout data_;
};
