// ref_shim_gpu.cu — TEST INFRASTRUCTURE.  C-ABI window onto the reference's own GPU launchers
// (src/include/kernel.hpp:8-17), recompiled UNMODIFIED for sm_100a from where they lie under
// $(REF) by oracle/Makefile into oracle/_ref/libspmv_ref_gpu.so (wsp_sm.cu is left out: it
// does not link — 64 KB of static shared memory).
//
// Purpose: on a GPU box, tests compare (a) this library's kernels and (b) the oracle's
// gpu_order emulations against what the reference's kernels actually produce on a B200, and
// bench.py reports the reference kernels' own device time next to ours.  Never part of the
// product path.
#include <cstring>
#include <iostream>
#include <sstream>
#include <string>
#include "kernel.hpp"

extern "C" {

// which: 0 cublas, 1 wsp, 2 asp, 3 awsp, 4 awsp_ref, 5 csr_naive, 6 csr_tiling, 7 naive, 8 tiling.
// Returns the milliseconds the reference's TIME_KERNEL macro printed (kernel.hpp:31-48: one
// cold launch bracketed by cudaEvents), or a negative value if nothing was printed
// (e.g. unknown `version`: no launch, wsp.cu:187-188).
float ref_gpu_gemv(int which, int version, int M, int N, const float *A_c, const float *x_c,
                   float *y)
{
    float *A = const_cast<float *>(A_c);
    float *x = const_cast<float *>(x_c);
    std::ostringstream cap;
    std::streambuf *old = std::cout.rdbuf(cap.rdbuf());
    switch (which) {
    case 0: cublas_gemv_gpu(M, N, A, x, y); break;
    case 1: wsp_gemv_gpu(M, N, A, x, y, version); break;
    case 2: asp_gemv_gpu(M, N, A, x, y, version); break;
    case 3: awsp_gemv_gpu(M, N, A, x, y, version); break;
    case 4: awsp_ref_gemv_gpu(M, N, A, x, y); break;
    case 5: csr_naive_gemv_gpu(M, N, A, x, y); break;
    case 6: csr_tiling_gemv_gpu(M, N, A, x, y); break;
    case 7: naive_gemv_gpu(M, N, A, x, y); break;
    case 8: tiling_gemv_gpu(M, N, A, x, y); break;
    default: break;
    }
    std::cout.rdbuf(old);
    const std::string s = cap.str();
    size_t p = s.rfind(" took ");
    if (p == std::string::npos) return -1.0f;
    return std::stof(s.substr(p + 6));
}

} // extern "C"
