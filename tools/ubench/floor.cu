// floor.cu — what a kernel launch and a plain streaming read cost in the bench's own timing
// method (CUDA graph of back-to-back launches, cold L2 by rotation over > 2.5x L2 of buffers,
// CUDA events).  Gives the attainable floor for the small configs: a matrix of S bytes cannot
// be processed faster than "read S bytes once" takes here, whatever the SGEMV kernel does.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o floor.bin floor.cu && ./floor.bin
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__global__ void empty_kernel(float *out) { if (out == nullptr && threadIdx.x == 9999) out[0] = 0.f; }

// every thread streams float4s with `unroll` independent loads in flight and folds them into one
// value (stored only if it is NaN-free-impossible, so the loads cannot be dropped)
template <int U>
__global__ void __launch_bounds__(256) read_kernel(const float4 *__restrict__ p, size_t n4, float *out)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    float acc = 0.f;
    for (; i + (U - 1) * stride < n4; i += U * stride) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; u++)
            asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                         : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w) : "l"(p + i + u * stride));
#pragma unroll
        for (int u = 0; u < U; u++) acc += v[u].x + v[u].y + v[u].z + v[u].w;
    }
    for (; i < n4; i += stride) { const float4 v = p[i]; acc += v.x + v.y + v.z + v.w; }
    if (acc == 123.456f) out[0] = acc;
}

static float time_graph(cudaStream_t st, int launches, int reps, void (*enqueue)(int, cudaStream_t, void *), void *ctx)
{
    cudaGraph_t g; cudaGraphExec_t ge;
    CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    for (int i = 0; i < launches; i++) enqueue(i, st, ctx);
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

struct ReadCtx { std::vector<float4 *> bufs; size_t n4; float *out; int grid; };

static void enq_empty(int, cudaStream_t st, void *ctx) { empty_kernel<<<148, 128, 0, st>>>((float *)ctx); }
static void enq_read(int i, cudaStream_t st, void *c)
{
    ReadCtx *r = (ReadCtx *)c;
    read_kernel<8><<<r->grid, 256, 0, st>>>(r->bufs[i % r->bufs.size()], r->n4, r->out);
}

int main()
{
    cudaStream_t st;
    CK(cudaStreamCreate(&st));
    float *out;
    CK(cudaMalloc(&out, 64));
    printf("empty kernel, graph of 100 launches: %.2f us per launch\n", time_graph(st, 100, 20, enq_empty, out));
    const double sizes_mb[] = {6.7, 13.5, 21.6, 33.6, 53.0, 70.6, 106.0, 115.3, 350.0};
    for (double mb : sizes_mb) {
        ReadCtx c;
        c.n4 = (size_t)(mb * 1e6 / 16);
        c.out = out;
        const int copies = (int)(320e6 / (mb * 1e6)) + 2;
        for (int k = 0; k < copies; k++) {
            float4 *p;
            CK(cudaMalloc(&p, c.n4 * 16));
            CK(cudaMemset(p, 0, c.n4 * 16));
            c.bufs.push_back(p);
        }
        float best = 1e9f; int best_grid = 0;
        for (int per_sm : {2, 4, 8}) {
            c.grid = 148 * per_sm;
            const float us = time_graph(st, 100, 5, enq_read, &c);
            if (us < best) { best = us; best_grid = c.grid; }
        }
        printf("read %7.1f MB (cold L2, %d copies): %7.2f us per call = %7.1f GB/s  (grid %d)\n", mb, copies, best,
               mb * 1e6 / (best * 1e-6) / 1e9, best_grid);
        for (float4 *p : c.bufs) cudaFree(p);
    }
    return 0;
}
