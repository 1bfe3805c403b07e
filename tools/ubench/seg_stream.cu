// Microbenchmark (development): ceiling of the AWSP data movement.  Every warp pulls every other
// row segment (values piece + index piece, two arrays) of its own region through a cp.async
// ring, no compute.  Sweeps piece size, warps per SM and ring depth.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../spmv_test_b200/csrc/common.cuh"
using namespace spmv;

template <int STG>
__global__ void __launch_bounds__(256) pull(const char *__restrict__ va, const char *__restrict__ ia, size_t span,
                                            int vbytes, int ibytes, int segs_per_warp, float *out)
{
    extern __shared__ __align__(128) unsigned char sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    unsigned char *ring = sm + (size_t)warp * STG * 768;
    const size_t wid = (size_t)blockIdx.x * nw + warp;
    const char *vb = va + (wid * (size_t)segs_per_warp * 2 * vbytes) % span;
    const char *ib = ia + (wid * (size_t)segs_per_warp * 2 * ibytes) % (span / 2);
    const int chunks_per_seg = (vbytes + 511) / 512;
    const int total = segs_per_warp * chunks_per_seg;
    float acc = 0.f;
    auto issue = [&](int i) {
        if (i < total) {
            const int seg = i / chunks_per_seg, c = i % chunks_per_seg, s = i % STG;
            const int vo = c * 512 + lane * 16, io = c * 256 + lane * 8;
            if (vo < vbytes) cp_async16(ring + s * 768 + lane * 16, vb + (size_t)seg * 2 * vbytes + vo);
            if (ibytes && io < ibytes) cp_async8(ring + s * 768 + 512 + lane * 8, ib + (size_t)seg * 2 * ibytes + io);
        }
        cp_async_commit();
    };
    for (int s = 0; s < STG; s++) issue(s);
    for (int i = 0; i < total; i++) {
        cp_async_wait<STG - 1>();
        acc += reinterpret_cast<float4 *>(ring + (i % STG) * 768)[lane].x;
        __syncwarp();
        issue(i + STG);
    }
    if (acc == 123.456f) out[0] = acc;
}

template <int STG> void run(const char *va, const char *ia, size_t span, int vb, int ib, int nw, int cps, float *out)
{
    const int ctas = 148 * cps;
    const int segs = (int)(((size_t)64 << 20) / (vb + ib) / (ctas * nw));
    const int smem = nw * STG * 768;
    cudaFuncSetAttribute(pull<STG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms = 0;
    for (int it = 0; it < 3; it++) {
        cudaEventRecord(e0);
        pull<STG><<<ctas, nw * 32, smem>>>(va, ia, span, vb, ib, segs, out);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
    }
    const double bytes = (double)segs * (vb + ib) * ctas * nw;
    printf("vals %5d B idx %4d B  warps/SM %2d  stages %2d  in-flight/SM %4d KB  %7.1f us  %7.1f GB/s (%s)\n", vb, ib,
           nw * cps, STG, nw * cps * STG * 768 / 1024, ms * 1e3, bytes / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    const size_t span = (size_t)1 << 30;
    char *va, *ia; float *out;
    cudaMalloc(&va, span + (1 << 22)); cudaMalloc(&ia, span / 2 + (1 << 22)); cudaMalloc(&out, 4);
    cudaMemset(va, 1, span); cudaMemset(ia, 1, span / 2);
    struct { int vb, ib; } shapes[] = {{320, 80}, {1232, 616}, {2464, 1232}, {4928, 2464}, {1232, 0}, {16384, 8192}};
    for (auto sh : shapes)
        for (int cps : {2, 4})
            for (int nw : {4, 8}) {
                if (nw * cps > 32) continue;
                run<8>(va, ia, span, sh.vb, sh.ib, nw, cps, out);
                run<16>(va, ia, span, sh.vb, sh.ib, nw, cps, out);
            }
    return 0;
}
