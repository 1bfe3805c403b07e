// awsp.hpp — drop-in for the reference's AWSPMatrix (src/include/awsp.hpp:4-18): per-tile row
// bitmaps + tile-padded values, public nz_bk_max_ (awsp.cpp:3-49).
#pragma once
#include <cstdint>
#include <vector>

#include "ref_layout.hpp"

class AWSPMatrix {
public:
    AWSPMatrix(int M, int N, float *matrix)
    {
        data_.Build(4, M, N, matrix);
        nz_bk_max_ = data_.aux[0];
    }

    auto GetBitmaps() -> uint32_t * { return data_.words.data(); }
    auto GetValues() -> float * { return data_.reals.data(); }
    auto BitmapsSize() -> int { return static_cast<int>(data_.words.size()); }
    auto ValuesSize() -> int { return static_cast<int>(data_.reals.size()); }

    int nz_bk_max_;

private:
    spmv_host::PackedLayout data_;
};
