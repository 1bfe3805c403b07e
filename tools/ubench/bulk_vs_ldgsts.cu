// Microbenchmark (development): how fast can one SM pull many small contiguous segments from
// HBM into shared memory — 1-D bulk async copies (TMA engine, one elected lane per warp) versus
// per-lane cp.async (LDGSTS)?  Decides how the panel kernel should stage its row segments.
//   ./bulk_vs_ldgsts
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../spmv_test_b200/csrc/common.cuh"
using namespace spmv;

constexpr int kStg = 8;

template <int MODE>   // 0: bulk copy per segment, 1: cp.async 16 B per lane
__global__ void __launch_bounds__(256) pull(const char *__restrict__ src, size_t span, int seg_bytes, int segs_per_warp,
                                            float *out)
{
    extern __shared__ __align__(128) unsigned char sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int slot = (seg_bytes + 127) / 128 * 128;
    unsigned char *ring = sm + (size_t)warp * (kStg * slot + 128);
    uint64_t *bars = reinterpret_cast<uint64_t *>(ring + kStg * slot);
    if (MODE == 0 && lane == 0) for (int s = 0; s < kStg; s++) mbar_init(&bars[s], 1);
    if (MODE == 0) { fence_mbar_init(); }
    __syncwarp();
    const size_t wid = (size_t)blockIdx.x * nw + warp;
    // every warp walks its own region with a stride of 2 segments (every other segment skipped)
    const char *base = src + (wid * (size_t)segs_per_warp * 2 * seg_bytes) % span;
    float acc = 0.f;
    auto issue = [&](int i, int s) {
        if (i >= segs_per_warp) { if (MODE == 1) cp_async_commit(); return; }
        const char *p = base + (size_t)i * 2 * seg_bytes;
        if (MODE == 0) {
            if (lane == 0) { mbar_expect_tx(&bars[s], seg_bytes); bulk_g2s(ring + s * slot, p, seg_bytes, &bars[s]); }
        } else {
            for (int o = lane * 16; o < seg_bytes; o += 512) cp_async16(ring + s * slot + o, p + o);
            cp_async_commit();
        }
    };
    for (int s = 0; s < kStg; s++) issue(s, s);
    for (int i = 0; i < segs_per_warp; i++) {
        const int s = i % kStg;
        if (MODE == 0) mbar_wait(&bars[s], (i / kStg) & 1); else cp_async_wait<kStg - 1>();
        if (lane * 16 < seg_bytes) acc += reinterpret_cast<float4 *>(ring + s * slot)[lane].x;
        __syncwarp();
        issue(i + kStg, s);
    }
    if (acc == 123.456f) out[0] = acc;
}

int main()
{
    const size_t span = (size_t)1 << 30;
    char *src; float *out;
    cudaMalloc(&src, span + (1 << 20)); cudaMemset(src, 1, span); cudaMalloc(&out, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int seg : {256, 400, 512, 1024, 2048})
        for (int nw : {4, 8})
            for (int mode = 0; mode < 2; mode++) {
                const int ctas = 148 * (nw == 4 ? 8 : 4);
                const int segs = (int)((size_t)96 << 20) / seg / (ctas * nw);     // ~96 MB per launch
                const int smem = nw * (kStg * ((seg + 127) / 128 * 128) + 128);
                auto k = mode == 0 ? pull<0> : pull<1>;
                cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
                for (int it = 0; it < 3; it++) {
                    cudaEventRecord(e0);
                    k<<<ctas, nw * 32, smem>>>(src, span, seg, segs, out);
                    cudaEventRecord(e1);
                    cudaEventSynchronize(e1);
                }
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                const double bytes = (double)segs * seg * ctas * nw;
                printf("seg %4d B  warps/CTA %d  %s  %7.1f us  %7.1f GB/s  (%s)\n", seg, nw, mode ? "cp.async" : "bulk    ",
                       ms * 1e3, bytes / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
            }
    return 0;
}
