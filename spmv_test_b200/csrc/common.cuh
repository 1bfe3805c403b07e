// common.cuh — device-side building blocks shared by the sm_100a SGEMV kernels.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace spmv {

// ---- development-only timeline (tools/trace_build.sh builds a -DSPMV_TRACE library) -------
#ifdef SPMV_TRACE
constexpr int kTraceSlots = 10;
constexpr int kTraceWarps = 1 << 16;
extern __device__ unsigned long long g_trace[kTraceWarps * kTraceSlots];
__device__ __forceinline__ void trace_stamp(int warp_global, int slot)
{
    if ((threadIdx.x & 31) == 0 && warp_global < kTraceWarps) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_trace[(size_t)warp_global * kTraceSlots + slot] = t;
    }
}
__device__ __forceinline__ void trace_smid(int warp_global, int slot)
{
    if ((threadIdx.x & 31) == 0 && warp_global < kTraceWarps) {
        unsigned id;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(id));
        g_trace[(size_t)warp_global * kTraceSlots + slot] = id;
    }
}
#define SPMV_STAMP(w, s) ::spmv::trace_stamp((w), (s))
#define SPMV_STAMP_SMID(w, s) ::spmv::trace_smid((w), (s))
#else
#define SPMV_STAMP(w, s) ((void)0)
#define SPMV_STAMP_SMID(w, s) ((void)0)
#endif

// ---- destinations of y ------------------------------------------------------------------------
// Single GPU: one pointer.  Column-sharded multi-GPU: the kernel's final
// stores go straight into every rank's copy of the full y — one multimem.st through the NVSwitch
// multicast address when there is one (NVLS), else one peer store per rank over NVLink — so the
// all-gather is fused into the epilogue and no separate collective kernel runs.
constexpr int kMaxYDst = 8;
struct YDst {
    float *p[kMaxYDst];      // base pointers, already offset to this rank's slice
    float *mc;               // multicast alias of the same slice (or null)
    int n;
    int act;                 // activation fused into the store: 0 none, 1 ReLU (spmv_run_act)
};
// ReLU as `v < 0 ? 0 : v`: NaN stays NaN; the next SGEMV's `x != 0` test then skips the zeros
__device__ __forceinline__ float y_act(const YDst &d, float v) { return (d.act == 1 && v < 0.0f) ? 0.0f : v; }
__device__ __forceinline__ void y_store(const YDst &d, size_t i, float v)
{
    v = y_act(d, v);
    if (d.mc) {
        asm volatile("multimem.st.weak.global.b32 [%0], %1;" ::"l"(d.mc + i), "r"(__float_as_uint(v)) : "memory");
    } else {
#pragma unroll 1
        for (int k = 0; k < d.n; k++) d.p[k][i] = v;
    }
}
__device__ __forceinline__ void y_store4(const YDst &d, size_t i4, float4 v)   // i4: index in float4 units
{
    v.x = y_act(d, v.x); v.y = y_act(d, v.y); v.z = y_act(d, v.z); v.w = y_act(d, v.w);
    if (d.mc) {
        unsigned long long lo, hi;
        asm("mov.b64 %0, {%1, %2};" : "=l"(lo) : "r"(__float_as_uint(v.x)), "r"(__float_as_uint(v.y)));
        asm("mov.b64 %0, {%1, %2};" : "=l"(hi) : "r"(__float_as_uint(v.z)), "r"(__float_as_uint(v.w)));
        asm volatile("multimem.st.weak.global.b64 [%0], %1;" ::"l"(d.mc + i4 * 4), "l"(lo) : "memory");
        asm volatile("multimem.st.weak.global.b64 [%0], %1;" ::"l"(d.mc + i4 * 4 + 2), "l"(hi) : "memory");
    } else {
#pragma unroll 1
        for (int k = 0; k < d.n; k++) reinterpret_cast<float4 *>(d.p[k])[i4] = v;
    }
}

// programmatic dependent launch (plan.hpp: launch_k): the kernel may have been placed before the
// previous grid in the stream has flushed; wait for it before touching anything it may have written
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

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
__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gmem_src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src)
                 : "memory");
}
// zero-fill forms: copy when `valid`, otherwise write zeros (src-size 0; src must still be a
// legal address).  Lets a partial chunk stay branch-free.
__device__ __forceinline__ void cp_async16_zfill(void *smem_dst, const void *gmem_src, bool valid)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src),
                 "r"(valid ? 16 : 0)
                 : "memory");
}
__device__ __forceinline__ void cp_async4_zfill(void *smem_dst, const void *gmem_src, bool valid)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src),
                 "r"(valid ? 4 : 0)
                 : "memory");
}
__device__ __forceinline__ void cp_async8_zfill(void *smem_dst, const void *gmem_src, bool valid)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src),
                 "r"(valid ? 8 : 0)
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
// Several CTAs have each written one partial row (`width` floats) for an output tile.  The
// last one to arrive (integer ticket, not a float atomic) adds the rows, so the result does
// not depend on which CTA happens to be last: with V = width/4 float4 lanes per row and
// K = blockDim/V thread groups, group k adds rows k, k+K, k+2K, ... in ascending order
// (kRedBatch independent 128-bit loads in flight), then the K group sums are added in group
// order.  `row(j)` returns the j-th partial row (16-byte aligned); `scratch` needs blockDim.x
// float4 of shared memory that is free by now.  Returns true in the CTA that performed the
// reduction.  All threads of the CTA must call it.
constexpr int kRedBatch = 8;

__device__ __forceinline__ float4 f4_add(float4 a, float4 b)
{
    return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}

template <class RowFn>
__device__ __forceinline__ bool split_reduce_rows(const YDst &yd, size_t tile_off, RowFn row, unsigned *__restrict__ ticket,
                                                  int rows, int width, int n_valid, int *smem_flag, float4 *scratch)
{
    __threadfence();      // publish this CTA's partial row
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = atomicAdd(ticket, 1u);
        int last = (t == (unsigned)(rows - 1));
        if (last) *ticket = 0; // re-arm for the next call / graph replay
        *smem_flag = last;
    }
    __syncthreads();
    if (!*smem_flag) return false;
    __threadfence();      // acquire the other CTAs' partial rows

    const int T = blockDim.x, V = width >> 2;
    auto sum_rows = [&](int v, int first, int step) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int s = first; s < rows; s += step * kRedBatch) {
            float4 t[kRedBatch];
#pragma unroll
            for (int u = 0; u < kRedBatch; u++) {
                const int su = s + u * step;
                t[u] = su < rows ? __ldcg(reinterpret_cast<const float4 *>(row(su)) + v) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < kRedBatch; u++) acc = f4_add(acc, t[u]);
        }
        return acc;
    };
    const size_t out4 = tile_off >> 2;                    // tile offsets are multiples of 4 floats
    if (T >= 2 * V) {
        const int K = T / V, k = threadIdx.x / V, v = threadIdx.x - k * V;
        if (k < K) scratch[threadIdx.x] = sum_rows(v, k, K);
        __syncthreads();
        if (k == 0 && v * 4 < n_valid) {
            float4 acc = scratch[v];
            for (int j = 1; j < K; j++) acc = f4_add(acc, scratch[j * V + v]);
            y_store4(yd, out4 + v, acc);
        }
    } else {
        for (int v = threadIdx.x; v * 4 < n_valid; v += T) y_store4(yd, out4 + v, sum_rows(v, 0, 1));
    }
    return true;
}

// regular layout: partial[split][tile*width ..], split stride `split_stride` floats
__device__ __forceinline__ bool split_reduce_finish(const YDst &yd, const float *__restrict__ partial,
                                                     unsigned *__restrict__ tickets, int tile, int splits,
                                                     int width, int n_valid, size_t split_stride,
                                                     int *smem_flag, float4 *scratch)
{
    const float *base = partial + (size_t)tile * width;
    return split_reduce_rows(yd, (size_t)tile * width, [&](int j) { return base + (size_t)j * split_stride; },
                             &tickets[tile], splits, width, n_valid, smem_flag, scratch);
}

} // namespace spmv
