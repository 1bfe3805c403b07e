// ref_layout.hpp — storage shared by the drop-in format classes.
//
// The reference keeps one std::vector per stream inside each format class and hands out raw
// pointers into them (matrix_csr.hpp:4-25, tcsr.hpp:4-23, wsp.hpp:4-20, asp.hpp:3-11,
// awsp.hpp:4-18, awsp_ref.hpp:4-18).  Here all six classes are thin views over one packed
// record produced by spmv_ref_pack() (include/spmv_b200.h), the bit-exact re-implementation
// of the reference packers inside libspmv_b200.so.
#pragma once
#include <cstdint>
#include <vector>

namespace spmv_host {

struct PackedLayout {
    std::vector<int> ints_a;          // row_pointers | blk_idx | warp_nz_offset
    std::vector<int> ints_b;          // col_indices
    std::vector<uint32_t> words;      // bitmaps
    std::vector<float> reals;         // values
    int aux[4] = {0, 0, 0, 0};

    // layout: spmv_layout_t.  Exits like the reference's CUDA_CHECK on failure.
    void Build(int layout, int m, int n, const float *dense);
};

} // namespace spmv_host
