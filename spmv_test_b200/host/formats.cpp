// formats.cpp — the drop-in format classes' storage, filled by the library's bit-exact
// re-implementation of the reference packers (spmv_ref_pack, csrc/pack_host.cpp).
#include <cstdio>
#include <cstdlib>

#include "matrix_csr.hpp"
#include "ref_layout.hpp"
#include "spmv_b200.h"

namespace spmv_host {

void PackedLayout::Build(int layout, int m, int n, const float *dense)
{
    spmv_ref_packed_t raw;
    const int rc = spmv_ref_pack(layout, m, n, dense, &raw);
    if (rc != SPMV_OK) {
        // the reference would have asserted (tester.cpp:9-10) or overflowed its int indices
        fprintf(stderr, "spmv format error: %s\n", spmv_last_error());
        exit(EXIT_FAILURE);
    }
    if (raw.i32_a) ints_a.assign(raw.i32_a, raw.i32_a + raw.n_i32_a);
    if (raw.i32_b) ints_b.assign(raw.i32_b, raw.i32_b + raw.n_i32_b);
    if (raw.u32) words.assign(raw.u32, raw.u32 + raw.n_u32);
    if (raw.f32) reals.assign(raw.f32, raw.f32 + raw.n_f32);
    for (int k = 0; k < 4; k++) aux[k] = raw.aux[k];
    spmv_ref_packed_free(&raw);
}

} // namespace spmv_host

void CSRMatrix::PrintCSR()
{
    std::cout << "CSR of A^T: " << cols_ << " lists over " << rows_ << " rows, " << ValuesSize() << " non-zeros\n";
    const int show = RowPtrsSize() < 8 ? RowPtrsSize() : 8;
    for (int i = 0; i < show; i++) {
        const int b = GetRowPtrs()[i], e = (i + 1 < RowPtrsSize()) ? GetRowPtrs()[i + 1] : ValuesSize();
        std::cout << "  y[" << i << "]:";
        for (int k = b; k < e && k < b + 8; k++) std::cout << " (" << GetColIdxs()[k] << ", " << GetValues()[k] << ")";
        if (e - b > 8) std::cout << " ...";
        std::cout << "\n";
    }
}
