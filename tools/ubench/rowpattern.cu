// rowpattern.cu — is the asp kernel's ACCESS PATTERN able to stream at the plain-read rate?
// Reads every second row of a dense row-major matrix (what asp does with half of x non-zero),
// CTA = (column tile, row range) with the tile `piece` bytes wide, nothing but loads and adds:
// no compaction, no shared memory, no split reduction.  Timed like bench.py (graph of
// back-to-back launches, rotation over copies > 2.5x L2).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o rowpattern.bin rowpattern.cu && ./rowpattern.bin
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

// thread t of a CTA owns float4 column t of the tile; U rows in flight per thread
template <int U>
__global__ void rows_kernel(const float4 *__restrict__ A, long long ld4, int M, int rows_per_cta, int row_step, float *out)
{
    const int tile = blockIdx.x, split = blockIdx.y;
    const float4 *p = A + (size_t)tile * blockDim.x + threadIdx.x;
    const int r0 = split * rows_per_cta, r1 = min(M, r0 + rows_per_cta);
    float acc = 0.f;
    int r = r0;
    for (; r + (U - 1) * row_step < r1; r += U * row_step) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; u++)
            asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                         : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w) : "l"(p + (size_t)(r + u * row_step) * ld4));
#pragma unroll
        for (int u = 0; u < U; u++) acc += v[u].x + v[u].y + v[u].z + v[u].w;
    }
    for (; r < r1; r += row_step) { const float4 v = p[(size_t)r * ld4]; acc += v.x + v.y + v.z + v.w; }
    if (acc == 123.456f) out[0] = acc;
}

struct Ctx { std::vector<float4 *> bufs; long long ld4; int M, rows_per_cta, row_step, threads; dim3 grid; float *out; int U; };

static void enq(int i, cudaStream_t st, void *c_)
{
    Ctx *c = (Ctx *)c_;
    const float4 *A = c->bufs[i % c->bufs.size()];
    if (c->U == 8) rows_kernel<8><<<c->grid, c->threads, 0, st>>>(A, c->ld4, c->M, c->rows_per_cta, c->row_step, c->out);
    else rows_kernel<16><<<c->grid, c->threads, 0, st>>>(A, c->ld4, c->M, c->rows_per_cta, c->row_step, c->out);
}

static float time_graph(cudaStream_t st, int launches, int reps, void *ctx)
{
    cudaGraph_t g; cudaGraphExec_t ge;
    CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    for (int i = 0; i < launches; i++) enq(i, st, ctx);
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
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    cudaGraphExecDestroy(ge); cudaGraphDestroy(g);
    return ms * 1e3f / (launches * reps);
}

int main()
{
    cudaStream_t st;
    CK(cudaStreamCreate(&st));
    float *out;
    CK(cudaMalloc(&out, 64));
    struct Shape { int M, N; } shapes[] = {{4096, 14336}, {4096, 4096}, {4096, 14336 + 32}, {4096, 4096 + 32}};
    for (Shape s : shapes) {
        Ctx c;
        c.M = s.M; c.ld4 = s.N / 4; c.out = out; c.row_step = 2;
        const size_t bytes = (size_t)s.M * s.N * 4;
        const int copies = (int)(320e6 / bytes) + 2;
        for (int k = 0; k < copies; k++) {
            float4 *p;
            CK(cudaMalloc(&p, bytes));
            CK(cudaMemset(p, 0, bytes));
            c.bufs.push_back(p);
        }
        const double read_mb = bytes / 2 / 1e6;
        for (int threads : {128, 256, 512}) {             // tile = threads * 16 bytes of a row
            const int tiles = (s.N / 4) / threads;
            if (tiles * threads * 4 != s.N / 4 * 4 && (s.N / 4) % threads) { /* ragged: skip the tail columns */ }
            for (int target : {148 * 2, 148 * 4, 148 * 8}) {
                int splits = target / (tiles > 0 ? tiles : 1);
                if (splits < 1) splits = 1;
                c.threads = threads; c.grid = dim3(tiles, splits);
                c.rows_per_cta = ((s.M + splits - 1) / splits + 1) / 2 * 2;
                for (int U : {8, 16}) {
                    c.U = U;
                    const float us = time_graph(st, 50, 4, &c);
                    printf("M=%d N=%d (row %d B) tile %5d B grid %3dx%-3d U=%2d: %7.2f us  %7.1f GB/s of %.1f MB\n", s.M, s.N, s.N * 4,
                           threads * 16, tiles, splits, U, us, read_mb * (double)(tiles * threads * 16) / (s.N * 4) / (us * 1e-6) / 1e3, read_mb);
                }
            }
        }
        for (float4 *p : c.bufs) cudaFree(p);
    }
    return 0;
}
