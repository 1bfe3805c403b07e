// compact.cu — activation compaction as an explicit pass.
//
// The reference tests x[j] != 0.0f inside its inner loops (asp.cu:23,64,76; awsp.cu:98,127,
// 228,258; awsp_ref.cu:52,96) and still loads every x and every bitmap word.  This pass writes
// the ascending list of rows with x != 0.0f (so -0.0f is dropped and NaN is kept, like the
// reference's compare), their values and the count.  Order preserving, deterministic,
// bit-exact with oracle/spmv_oracle.c: orc_compact_x.  The SGEMV kernels fuse the same
// ballot/popc compaction into their prologues; this entry point serves callers that want the
// list itself (and the tests that pin it).
#include "common.cuh"
#include "plan.hpp"

namespace spmv {

namespace {

constexpr int kCmpThreads = 256;
constexpr int kCmpWarps = kCmpThreads / 32;
constexpr int kCmpSteps = 8;                          // 32-row steps per warp
constexpr int kCmpChunk = kCmpThreads * kCmpSteps;    // 2048 rows per block pass
constexpr int64_t kCmpSingleBlockMax = 32 * 1024;

// Compacts rows [r0, r0 + kCmpChunk) starting at output position `out_base`; returns the
// number of rows kept (same value in every thread).
__device__ __forceinline__ int compact_chunk(const float *__restrict__ x, int64_t M, int64_t r0, int64_t out_base,
                                             int32_t *__restrict__ idx, float *__restrict__ val, int *wcnt,
                                             bool write)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float xr[kCmpSteps]; unsigned bal[kCmpSteps];
    int cnt = 0;
#pragma unroll
    for (int k = 0; k < kCmpSteps; k++) {
        const int64_t row = r0 + (int64_t)(warp * kCmpSteps + k) * 32 + lane;
        xr[k] = row < M ? __ldg(x + row) : 0.0f;
        bal[k] = __ballot_sync(kFull, xr[k] != 0.0f);
        cnt += __popc(bal[k]);
    }
    __syncthreads();
    if (lane == 0) wcnt[warp] = cnt;
    __syncthreads();
    int base = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kCmpWarps; w++) {
        const int c = wcnt[w];
        if (w < warp) base += c;
        total += c;
    }
    if (write) {
        const unsigned lt = (1u << lane) - 1u;
#pragma unroll
        for (int k = 0; k < kCmpSteps; k++) {
            if (xr[k] != 0.0f) {
                const int64_t pos = out_base + base + __popc(bal[k] & lt);
                idx[pos] = (int32_t)(r0 + (int64_t)(warp * kCmpSteps + k) * 32 + lane);
                val[pos] = xr[k];
            }
            base += __popc(bal[k]);
        }
    }
    return total;
}

__global__ void __launch_bounds__(kCmpThreads)
compact_single_block(const float *__restrict__ x, int64_t M, int32_t *__restrict__ idx, float *__restrict__ val,
                     int32_t *__restrict__ count)
{
    __shared__ int wcnt[kCmpWarps];
    int64_t out = 0;
    for (int64_t r0 = 0; r0 < M; r0 += kCmpChunk) out += compact_chunk(x, M, r0, out, idx, val, wcnt, true);
    if (threadIdx.x == 0) *count = (int32_t)out;
}

__global__ void __launch_bounds__(kCmpThreads)
compact_count(const float *__restrict__ x, int64_t M, uint32_t *__restrict__ block_cnt)
{
    __shared__ int wcnt[kCmpWarps];
    const int total = compact_chunk(x, M, (int64_t)blockIdx.x * kCmpChunk, 0, nullptr, nullptr, wcnt, false);
    if (threadIdx.x == 0) block_cnt[blockIdx.x] = (uint32_t)total;
}

// exclusive scan of the block counts, in place, by one CTA; also writes the total
__global__ void __launch_bounds__(1024)
compact_scan(uint32_t *__restrict__ block_cnt, int nblocks, int32_t *__restrict__ count)
{
    __shared__ uint32_t wsum[32];
    __shared__ uint32_t carry_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int b0 = 0; b0 < nblocks; b0 += 1024) {
        const int b = b0 + threadIdx.x;
        const uint32_t v = b < nblocks ? block_cnt[b] : 0u;
        uint32_t inc = (uint32_t)warp_incl_scan((int)v, lane);
        if (lane == 31) wsum[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            uint32_t s = wsum[lane];
            s = (uint32_t)warp_incl_scan((int)s, lane);
            wsum[lane] = s;
        }
        __syncthreads();
        const uint32_t carry = carry_s;
        const uint32_t before = carry + (warp ? wsum[warp - 1] : 0u) + inc - v;
        if (b < nblocks) block_cnt[b] = before;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + wsum[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) *count = (int32_t)carry_s;
}

__global__ void __launch_bounds__(kCmpThreads)
compact_write(const float *__restrict__ x, int64_t M, const uint32_t *__restrict__ block_off,
              int32_t *__restrict__ idx, float *__restrict__ val)
{
    __shared__ int wcnt[kCmpWarps];
    compact_chunk(x, M, (int64_t)blockIdx.x * kCmpChunk, (int64_t)block_off[blockIdx.x], idx, val, wcnt, true);
}

} // namespace

size_t compact_scratch_bytes(int64_t M)
{
    if (M <= kCmpSingleBlockMax) return 0;
    const int64_t blocks = (M + kCmpChunk - 1) / kCmpChunk;
    return (size_t)((blocks * 4 + 255) / 256 * 256);
}

int launch_compact(const float *d_x, int64_t M, int32_t *d_idx, float *d_val, int32_t *d_count,
                   void *d_scratch, size_t scratch_bytes, cudaStream_t st)
{
    if (M < 0 || M > INT32_MAX) return set_error(SPMV_ERR_SHAPE, "spmv_compact_x: M out of range");
    if (M <= kCmpSingleBlockMax) {
        compact_single_block<<<1, kCmpThreads, 0, st>>>(d_x, M, d_idx, d_val, d_count);
        SPMV_CUDA(cudaGetLastError());
        return SPMV_OK;
    }
    if (!d_scratch || scratch_bytes < compact_scratch_bytes(M))
        return set_error(SPMV_ERR_ARG, "spmv_compact_x: scratch of %zu bytes required", compact_scratch_bytes(M));
    const int blocks = (int)((M + kCmpChunk - 1) / kCmpChunk);
    uint32_t *cnt = reinterpret_cast<uint32_t *>(d_scratch);
    compact_count<<<blocks, kCmpThreads, 0, st>>>(d_x, M, cnt);
    compact_scan<<<1, 1024, 0, st>>>(cnt, blocks, d_count);
    compact_write<<<blocks, kCmpThreads, 0, st>>>(d_x, M, cnt, d_idx, d_val);
    SPMV_CUDA(cudaGetLastError());
    return SPMV_OK;
}

} // namespace spmv
