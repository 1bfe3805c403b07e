// common.cuh — device-side building blocks shared by the sm_100a SGEMV kernels.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace spmv {

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

// ---- streaming loads: 128-bit, read-only path, do not pollute L1 -------------------------
__device__ __forceinline__ float4 ldg_stream_f4(const float4 *p)
{
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint4 ldg_stream_u4(const uint4 *p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint2 ldg_stream_u2(const uint2 *p)
{
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];"
                 : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- Ampere-style async copy (LDGSTS): 16 B global -> shared, no register staging ----------
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src)
                 : "memory");
}
__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gmem_src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src)
                 : "memory");
}
__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gmem_src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// ---- mbarrier + 1-D bulk async copy (the TMA engine without a tensor map) -------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    }
}
// bytes must be a multiple of 16, both addresses 16-byte aligned.
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---- fixed-order warp reductions / scans (deterministic: no atomics anywhere) ---------------
__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) v += __shfl_xor_sync(kFull, v, s);
    return v;
}
__device__ __forceinline__ int warp_incl_scan(int v, int lane)
{
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
        int n = __shfl_up_sync(kFull, v, s);
        if (lane >= s) v += n;
    }
    return v;
}

// ---- cross-CTA fixed-order split reduction -------------------------------------------------
// Every CTA (tile, split) has written `width` partial sums to partial[split][tile*width ..].
// The last CTA to arrive for a tile (integer ticket, not a float atomic) adds the splits in
// ascending split order, so the result does not depend on which CTA happens to be last.
// Returns true in the CTA that performed the reduction.  All threads of the CTA must call it.
__device__ __forceinline__ bool split_reduce_finish(float *__restrict__ y, const float *__restrict__ partial,
                                                     unsigned *__restrict__ tickets, int tile, int splits,
                                                     int width, int n_valid, size_t split_stride,
                                                     int *smem_flag)
{
    __threadfence();      // publish this CTA's partials
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = atomicAdd(&tickets[tile], 1u);
        int last = (t == (unsigned)(splits - 1));
        if (last) tickets[tile] = 0; // re-arm for the next call / graph replay
        *smem_flag = last;
    }
    __syncthreads();
    if (!*smem_flag) return false;
    __threadfence();      // acquire the other CTAs' partials
    for (int c = threadIdx.x; c < width; c += blockDim.x) {
        if (c >= n_valid) break;
        const float *p = partial + (size_t)tile * width + c;
        float acc = 0.0f;
        for (int s = 0; s < splits; s++) acc += __ldcg(p + (size_t)s * split_stride);
        y[(size_t)tile * width + c] = acc;
    }
    return true;
}

} // namespace spmv
