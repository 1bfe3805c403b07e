// gather4.cu — can the TMA engine's row gather (cp.async.bulk.tensor.2d ... tile::gather4, sm_100) stream the asp
// kernel's ACCESS PATTERN faster than plain loads?  (SURVEY section 7 proposes it for asp; round 1's verdict asks for the
// measurement.)  Every second row of a dense row-major 4096 x 14336 fp32 matrix, CTA = (512-column tile, row range),
// exactly asp's decomposition on config 2, nothing but the loads and adds:
//   * gather4: one producer lane issues, per group of four active rows, two gather4 copies (4 rows x 256 columns = 4 KB
//     each) into a ring of STAGES 8 KB stages; full/empty mbarriers; four consumer warps read their 16 bytes per row back;
//   * registers: every thread loads its 16 bytes of U rows straight into registers (what asp.cu does for long lists).
// Timed like bench.py: a CUDA graph of back-to-back launches rotating over three copies of the matrix (705 MB > 2.5 x L2).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather4.bin gather4.cu && ./gather4.bin <box_rows 1|4> <stages>
// Every wait is bounded (a wrong tensor map makes the kernel trap, not hang), results are checked on the host.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int kM = 4096, kN = 14336, kTile = 512, kSplits = 8, kRowStep = 2;

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, int n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_expect(uint64_t *b, uint32_t tx) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(tx) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity)
{
    for (int spin = 0; spin < (1 << 20); spin++) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(s32(b)), "r"(parity) : "memory");
        if (ok) return;
    }
    __trap();                                             // never hang the box
}

template <int STAGES>
__global__ void __launch_bounds__(160)
gather4_kernel(const __grid_constant__ CUtensorMap tm, int rows_per_cta, float *out)
{
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full[STAGES], empty[STAGES];
    unsigned char *smem = smem_raw + ((1024u - (s32(smem_raw) & 1023u)) & 1023u);   // TMA destinations: generously aligned
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tile = blockIdx.x, split = blockIdx.y;
    const int r0 = split * rows_per_cta, ngroups = rows_per_cta / (4 * kRowStep);
    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp == 4) {
        if (lane == 0) {
            for (int i = 0; i < ngroups; i++) {
                const int s = i % STAGES;
                if (i >= STAGES) mbar_wait(&empty[s], ((i / STAGES) - 1) & 1);
                mbar_expect(&full[s], 8192u);
                const int r = r0 + i * 4 * kRowStep;
#pragma unroll
                for (int h = 0; h < 2; h++)
                    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                                 ::"r"(s32(smem + s * 8192 + h * 4096)), "l"(&tm), "r"(tile * kTile + h * 256), "r"(r), "r"(r + kRowStep),
                                 "r"(r + 2 * kRowStep), "r"(r + 3 * kRowStep), "r"(s32(&full[s])) : "memory");
            }
        }
        return;
    }
    float acc = 0.f;
    for (int i = 0; i < ngroups; i++) {
        const int s = i % STAGES;
        mbar_wait(&full[s], (i / STAGES) & 1);
        const float4 *base = reinterpret_cast<const float4 *>(smem + s * 8192 + (tid >> 6) * 4096) + (tid & 63);
#pragma unroll
        for (int q = 0; q < 4; q++) { const float4 v = base[q * 64]; acc += (v.x + v.y) + (v.z + v.w); }
        mbar_arrive(&empty[s]);
    }
    out[((size_t)split * gridDim.x + tile) * 128 + tid] = acc;
}

template <int U>
__global__ void __launch_bounds__(128)
regs_kernel(const float4 *__restrict__ A, long long ld4, int rows_per_cta, float *out)
{
    const int tile = blockIdx.x, split = blockIdx.y, tid = threadIdx.x;
    const float4 *p = A + (size_t)tile * 128 + tid;
    const int r0 = split * rows_per_cta, nrows = rows_per_cta / kRowStep;
    float acc = 0.f;
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; u++)
        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w)
                     : "l"(p + (size_t)(r0 + u * kRowStep) * ld4));
    for (int i0 = 0; i0 < nrows; i0 += U) {
#pragma unroll
        for (int u = 0; u < U; u++) {
            acc += (v[u].x + v[u].y) + (v[u].z + v[u].w);
            const int j = i0 + u + U;
            if (j < nrows)
                asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w)
                             : "l"(p + (size_t)(r0 + j * kRowStep) * ld4));
        }
    }
    out[((size_t)split * gridDim.x + tile) * 128 + tid] = acc;
}

typedef CUresult (*EncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <class F> static float time_graph(cudaStream_t st, int launches, int reps, F enq)
{
    cudaGraph_t g; cudaGraphExec_t ge;
    CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    for (int i = 0; i < launches; i++) enq(i);
    CK(cudaStreamEndCapture(st, &g));
    CK(cudaGraphInstantiate(&ge, g, 0));
    CK(cudaGraphLaunch(ge, st));
    CK(cudaStreamSynchronize(st));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0, st));
    for (int r = 0; r < reps; r++) CK(cudaGraphLaunch(ge, st));
    CK(cudaEventRecord(e1, st));
    CK(cudaStreamSynchronize(st));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms * 1000.f / (float)(launches * reps);
}

static bool check(const std::vector<float> &hA, const float *d_out, const char *what)
{
    const int tiles = kN / kTile, rows_per_cta = kM / kSplits;
    std::vector<float> out((size_t)kSplits * tiles * 128);
    CK(cudaMemcpy(out.data(), d_out, out.size() * sizeof(float), cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int split = 0; split < kSplits; split += 3)
        for (int tile = 0; tile < tiles; tile += 5)
            for (int t = 0; t < 128; t += 17) {
                double s = 0;
                for (int r = split * rows_per_cta; r < (split + 1) * rows_per_cta; r += kRowStep)
                    for (int c = 0; c < 4; c++) s += hA[(size_t)r * kN + (size_t)tile * kTile + t * 4 + c];
                if ((float)s != out[((size_t)split * tiles + tile) * 128 + t]) bad++;
            }
    printf("%s: results %s\n", what, bad ? "WRONG" : "ok");
    return bad == 0;
}

int main(int argc, char **argv)
{
    const int box_rows = argc > 1 ? atoi(argv[1]) : 1, stages = argc > 2 ? atoi(argv[2]) : 8;
    const int tiles = kN / kTile, rows_per_cta = kM / kSplits, copies = 3;
    std::vector<float> hA((size_t)kM * kN);
    uint32_t lcg = 12345u;
    for (float &v : hA) { lcg = lcg * 1664525u + 1013904223u; v = (float)((lcg >> 24) & 7u); }   // small integers: exact sums
    std::vector<float *> dA(copies);
    for (float *&p : dA) { CK(cudaMalloc(&p, hA.size() * sizeof(float))); CK(cudaMemcpy(p, hA.data(), hA.size() * sizeof(float), cudaMemcpyHostToDevice)); }
    float *d_out;
    CK(cudaMalloc(&d_out, (size_t)kSplits * tiles * 128 * sizeof(float)));
    cudaStream_t st;
    CK(cudaStreamCreate(&st));
    const double mb = (double)kM / kRowStep * kN * 4 / 1e6;
    const dim3 grid(tiles, kSplits);

    CK(cudaMemset(d_out, 0, (size_t)kSplits * tiles * 128 * sizeof(float)));
    regs_kernel<32><<<grid, 128, 0, st>>>(reinterpret_cast<const float4 *>(dA[0]), kN / 4, rows_per_cta, d_out);
    CK(cudaStreamSynchronize(st));
    check(hA, d_out, "registers, 32 rows in flight");
    float us = time_graph(st, 12, 10, [&](int i) { regs_kernel<32><<<grid, 128, 0, st>>>(reinterpret_cast<const float4 *>(dA[i % copies]), kN / 4, rows_per_cta, d_out); });
    printf("registers U=32  grid %dx%d x128: %.2f us  %.0f GB/s\n", tiles, kSplits, us, mb / us * 1e3);
    us = time_graph(st, 12, 10, [&](int i) { regs_kernel<48><<<grid, 128, 0, st>>>(reinterpret_cast<const float4 *>(dA[i % copies]), kN / 4, rows_per_cta, d_out); });
    printf("registers U=48  grid %dx%d x128: %.2f us  %.0f GB/s\n", tiles, kSplits, us, mb / us * 1e3);

    EncodeTiled encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", reinterpret_cast<void **>(&encode), cudaEnableDefault, &qres));
    if (!encode || qres != cudaDriverEntryPointSuccess) { printf("cuTensorMapEncodeTiled not available\n"); return 1; }
    std::vector<CUtensorMap> tms(copies);
    for (int i = 0; i < copies; i++) {
        const cuuint64_t dims[2] = {(cuuint64_t)kN, (cuuint64_t)kM};
        const cuuint64_t strides[1] = {(cuuint64_t)kN * 4};
        const cuuint32_t box[2] = {256u, (cuuint32_t)box_rows};
        const cuuint32_t estr[2] = {1u, 1u};
        const CUresult r = encode(&tms[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dA[i], dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d (box rows %d)\n", (int)r, box_rows); return 1; }
    }
    auto launch = [&](int i) {
        const size_t smem = (size_t)stages * 8192 + 1024;
        if (stages == 4) gather4_kernel<4><<<grid, 160, smem, st>>>(tms[i % copies], rows_per_cta, d_out);
        else if (stages == 12) gather4_kernel<12><<<grid, 160, smem, st>>>(tms[i % copies], rows_per_cta, d_out);
        else gather4_kernel<8><<<grid, 160, smem, st>>>(tms[i % copies], rows_per_cta, d_out);
    };
    CK(cudaFuncSetAttribute(gather4_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 8192 + 1024));
    CK(cudaFuncSetAttribute(gather4_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 8192 + 1024));
    CK(cudaFuncSetAttribute(gather4_kernel<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, 12 * 8192 + 1024));
    CK(cudaMemset(d_out, 0, (size_t)kSplits * tiles * 128 * sizeof(float)));
    launch(0);
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    if (!check(hA, d_out, "gather4")) { printf("gather4 with box rows %d: wrong data, not timed\n", box_rows); return 2; }
    us = time_graph(st, 12, 10, launch);
    printf("gather4 box_rows=%d stages=%d  grid %dx%d x(128+32): %.2f us  %.0f GB/s  (%d KB in flight per CTA)\n", box_rows, stages, tiles, kSplits,
           us, mb / us * 1e3, stages * 8);
    return 0;
}
