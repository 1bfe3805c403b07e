// mg.cu — column-sharded execution behind the C-ABI (spmv_mg_*, include/spmv_b200.h).
//
// the outputs are independent, so A is cut into column slabs; a rank (one GPU, one
// process — or one of several devices of one process) owns one or more slabs as ordinary plans.
// One call runs the rank's plans back to back; every plan's epilogue stores its slice of y into
// EVERY rank's copy of the full y (peer stores over NVLink, or one multimem.st through the NVSwitch
// multicast alias), so the all-gather is fused into the kernels.  What is left of the collective
// is the arrival: a one-warp kernel at the end of the call publishes this rank's epoch into every
// peer's flag word (st.release.sys) and spins until all peers' words have reached the epoch
// (ld.acquire.sys) — no host round trip, no separate collective launch, CUDA-graph capturable.
// Two y buffers alternate (the epoch's parity picks one) so that a fast rank starting call k+1
// cannot overwrite the y of call k a slower peer is still reading; reaching call k+2 requires the
// peer's arrival for k+1, which is stream-ordered after that peer's readers of y(k).
//
// The shared block of a rank:   [ y0: Npad floats ][ y1: Npad floats ][ flags: 32 x u32 ]
// It is either allocated here (cudaMalloc; other processes map it through a CUDA IPC handle,
// other devices of the same process through peer access) or supplied by the caller (symmetric
// memory with a multicast alias, e.g. torch.distributed._symmetric_memory).
#include <algorithm>
#include <cstring>
#include <new>
#include <utility>
#include <vector>

#include "common.cuh"
#include "plan.hpp"

namespace spmv {

constexpr int kMgFlagWords = 32;
constexpr unsigned long long kMgTimeoutNs = 4000000000ull;   // a peer that never arrives: give up after 4 s

struct MgJoin {
    unsigned *peer_flags[kMaxYDst];   // rank r's flag array as mapped here
    unsigned *my_flags;
    unsigned *epoch;                  // device word: calls completed by this rank
    unsigned *status;                 // device word: 1 after a timed-out wait
    int rank, world;
};

__global__ void __launch_bounds__(32) mg_join_kernel(const MgJoin j)
{
    pdl_trigger();                                        // the next call's first kernel may come up and wait behind this one
    pdl_wait();                                           // every kernel of this call has finished and flushed
    const int t = threadIdx.x;
    unsigned e = 0;
    if (t == 0) {
        e = *j.epoch + 1u;
        *j.epoch = e;
        __threadfence_system();                           // this call's y stores before the flags
    }
    e = __shfl_sync(kFull, e, 0);
    if (t < j.world && t != j.rank)
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(j.peer_flags[t] + j.rank), "r"(e) : "memory");
    if (t < j.world && t != j.rank) {
        unsigned long long t0, t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        for (;;) {
            unsigned v;
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(j.my_flags + t) : "memory");
            if ((int)(v - e) >= 0) break;                 // the peer has completed call e (or a later one)
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > kMgTimeoutNs) { *j.status = 1u; break; }
            __nanosleep(64);
        }
    }
}

} // namespace spmv

using namespace spmv;

struct spmv_mg {
    int64_t M = 0, N = 0, Npad = 0;
    int rank = 0, world = 1, device = 0;
    bool own_block = false, ipc_opened = false, connected = false;
    char *block = nullptr;                    // this rank's shared block
    char *peer_block[kMaxYDst] = {};          // every rank's block as mapped here (peer_block[rank] == block)
    char *mc_block = nullptr;                 // multicast alias of the blocks, or null
    unsigned *epoch = nullptr, *status = nullptr;
    float *d_x = nullptr;                     // staging for run_host
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;       // run_host: a slab's columns travel to the host while the next slab computes
    std::vector<cudaEvent_t> part_done;       // one per part, recorded on `stream` behind the part's kernels
    uint64_t calls = 0;                       // host mirror of the epoch (picks the y buffer)
    struct Part { spmv_plan *plan; int64_t off; };
    std::vector<Part> parts;
};

static size_t mg_npad(int64_t N) { return (size_t)((N + 63) / 64 * 64); }

extern "C" {

size_t spmv_mg_block_bytes(int64_t N_total)
{
    if (N_total < 0) return 0;
    return 2 * mg_npad(N_total) * sizeof(float) + kMgFlagWords * sizeof(unsigned);
}

int spmv_mg_create(int64_t M, int64_t N_total, int rank, int world, void *local_block, spmv_mg_t **out)
{
    if (!out) return set_error(SPMV_ERR_ARG, "null output pointer");
    *out = nullptr;
    if (M < 0 || N_total < 0 || N_total % 4) return set_error(SPMV_ERR_SHAPE, "bad shape %lld x %lld (N must be a multiple of 4)", (long long)M, (long long)N_total);
    if (world < 1 || world > kMaxYDst || rank < 0 || rank >= world) return set_error(SPMV_ERR_ARG, "rank %d of %d (at most %d ranks)", rank, world, kMaxYDst);
    spmv_mg *g = new (std::nothrow) spmv_mg();
    if (!g) return set_error(SPMV_ERR_NOMEM, "out of host memory");
    g->M = M; g->N = N_total; g->Npad = (int64_t)mg_npad(N_total); g->rank = rank; g->world = world;
    auto fail = [&](int rc) { spmv_mg_destroy(g); return rc; };
    if (cudaGetDevice(&g->device) != cudaSuccess) { cudaGetLastError(); return fail(set_error(SPMV_ERR_CUDA, "no CUDA device (this library has no CPU path)")); }
    const size_t bytes = spmv_mg_block_bytes(N_total);
    cudaError_t e = cudaSuccess;
    if (local_block) g->block = reinterpret_cast<char *>(local_block);
    else { e = cudaMalloc(reinterpret_cast<void **>(&g->block), bytes); g->own_block = e == cudaSuccess; }
    if (e == cudaSuccess) e = cudaMemset(g->block, 0, bytes);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void **>(&g->epoch), 2 * sizeof(unsigned));
    if (e == cudaSuccess) e = cudaMemset(g->epoch, 0, 2 * sizeof(unsigned));
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void **>(&g->d_x), ((size_t)M + 4) * sizeof(float));
    if (e == cudaSuccess) e = cudaMemset(g->d_x, 0, ((size_t)M + 4) * sizeof(float));
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&g->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&g->copy_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) return fail(cuda_error(e, "spmv_mg_create"));
    g->status = g->epoch + 1;
    g->peer_block[rank] = g->block;
    g->connected = world == 1;
    *out = g;
    return SPMV_OK;
}

void spmv_mg_destroy(spmv_mg_t *g)
{
    if (!g) return;
    if (g->ipc_opened)
        for (int r = 0; r < g->world; r++)
            if (r != g->rank && g->peer_block[r]) cudaIpcCloseMemHandle(g->peer_block[r]);
    if (g->own_block && g->block) cudaFree(g->block);
    if (g->epoch) cudaFree(g->epoch);
    if (g->d_x) cudaFree(g->d_x);
    if (g->stream) cudaStreamDestroy(g->stream);
    if (g->copy_stream) cudaStreamDestroy(g->copy_stream);
    for (cudaEvent_t e : g->part_done) cudaEventDestroy(e);
    cudaGetLastError();
    delete g;
}

int spmv_mg_ipc_handle(spmv_mg_t *g, void *handle64)
{
    if (!g || !handle64) return set_error(SPMV_ERR_ARG, "null argument");
    if (!g->own_block) return set_error(SPMV_ERR_ARG, "the shared block was supplied by the caller: exchange it the way it was allocated");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
    cudaIpcMemHandle_t h;
    SPMV_CUDA(cudaIpcGetMemHandle(&h, g->block));
    std::memcpy(handle64, &h, 64);
    return SPMV_OK;
}

int spmv_mg_connect_ipc(spmv_mg_t *g, const void *handles)
{
    if (!g || (!handles && g->world > 1)) return set_error(SPMV_ERR_ARG, "null argument");
    if (g->connected && g->world > 1) return set_error(SPMV_ERR_ARG, "already connected");
    for (int r = 0; r < g->world; r++) {
        if (r == g->rank) continue;
        cudaIpcMemHandle_t h;
        std::memcpy(&h, reinterpret_cast<const char *>(handles) + (size_t)r * 64, 64);
        void *p = nullptr;
        SPMV_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        g->peer_block[r] = reinterpret_cast<char *>(p);
        g->ipc_opened = true;
    }
    g->connected = true;
    return SPMV_OK;
}

int spmv_mg_connect_ptrs(spmv_mg_t *g, void *const *blocks, void *mc_block)
{
    if (!g || (!blocks && g->world > 1)) return set_error(SPMV_ERR_ARG, "null argument");
    for (int r = 0; r < g->world; r++) {
        if (r == g->rank) continue;
        if (!blocks[r]) return set_error(SPMV_ERR_ARG, "block of rank %d is null", r);
        g->peer_block[r] = reinterpret_cast<char *>(blocks[r]);
    }
    g->mc_block = reinterpret_cast<char *>(mc_block);
    g->connected = true;
    return SPMV_OK;
}

int spmv_mg_add_plan(spmv_mg_t *g, spmv_plan_t *plan, int64_t col_offset)
{
    if (!g || !plan) return set_error(SPMV_ERR_ARG, "null argument");
    if (plan->M != g->M) return set_error(SPMV_ERR_SHAPE, "plan has %lld rows, the group %lld", (long long)plan->M, (long long)g->M);
    if (col_offset < 0 || col_offset % 4 || col_offset + plan->N > g->N)
        return set_error(SPMV_ERR_SHAPE, "columns [%lld, %lld) do not fit the group's %lld (offset must be a multiple of 4)",
                         (long long)col_offset, (long long)(col_offset + plan->N), (long long)g->N);
    if (plan->device != g->device) return set_error(SPMV_ERR_ARG, "plan lives on device %d, the group on %d", plan->device, g->device);
    g->parts.push_back({plan, col_offset});
    return SPMV_OK;
}

} // extern "C"

// the step; with `mark` an event is recorded on the stream behind every part's kernels (run_host's copy-back pipeline)
static int mg_run_step(spmv_mg *g, const float *d_x, void *stream, const float **d_y, bool mark)
{
    if (!g) return set_error(SPMV_ERR_ARG, "null group");
    if (!g->connected) return set_error(SPMV_ERR_ARG, "spmv_mg_run before the ranks were connected");
    if (!d_x && g->M > 0) return set_error(SPMV_ERR_ARG, "null device vector");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const size_t ybase = (size_t)(g->calls & 1) * (size_t)g->Npad * sizeof(float);
    for (const spmv_mg::Part &pt : g->parts) {
        float *dst[kMaxYDst];
        for (int r = 0; r < g->world; r++) dst[r] = reinterpret_cast<float *>(g->peer_block[r] + ybase);
        float *mc = g->mc_block ? reinterpret_cast<float *>(g->mc_block + ybase) : nullptr;
        // own rank first is not required: the epilogue stores to every destination alike
        int rc = spmv_run_scatter(pt.plan, d_x, g->world, dst, mc, pt.off, st);
        if (rc) return rc;
        if (mark) SPMV_CUDA(cudaEventRecord(g->part_done[(size_t)(&pt - g->parts.data())], st));
    }
    if (g->world > 1) {
        MgJoin j{};
        const size_t foff = 2 * (size_t)g->Npad * sizeof(float);
        for (int r = 0; r < g->world; r++) j.peer_flags[r] = reinterpret_cast<unsigned *>(g->peer_block[r] + foff);
        j.my_flags = reinterpret_cast<unsigned *>(g->block + foff);
        j.epoch = g->epoch; j.status = g->status; j.rank = g->rank; j.world = g->world;
        SPMV_CUDA(launch_k(mg_join_kernel, dim3(1), dim3(32), 0, st, j));
    }
    if (d_y) *d_y = reinterpret_cast<const float *>(g->block + ybase);
    g->calls++;
    return SPMV_OK;
}

// Copy-back of run_host.  A part's own columns of y are final in this rank's block as soon as the part's
// kernels have run (the arrival only concerns the peers' columns), so when the requested range lies inside
// this rank's parts each part's columns leave on the copy stream behind the part's event while the later
// parts still compute; otherwise one copy behind the whole step.  Nothing is synchronised here.
static int mg_copy_back(spmv_mg *g, const float *dy, float *y, int64_t y_begin, int64_t y_count, bool marked)
{
    if (y_count <= 0) return SPMV_OK;
    if (!marked) {
        SPMV_CUDA(cudaMemcpyAsync(y, dy + y_begin, (size_t)y_count * sizeof(float), cudaMemcpyDeviceToHost, g->stream));
        return SPMV_OK;
    }
    for (size_t k = 0; k < g->parts.size(); k++) {
        const spmv_mg::Part &pt = g->parts[k];
        const int64_t a = std::max(y_begin, pt.off), b = std::min(y_begin + y_count, pt.off + pt.plan->N);
        if (b <= a) continue;
        SPMV_CUDA(cudaStreamWaitEvent(g->copy_stream, g->part_done[k], 0));
        SPMV_CUDA(cudaMemcpyAsync(y + (a - y_begin), dy + a, (size_t)(b - a) * sizeof(float), cudaMemcpyDeviceToHost, g->copy_stream));
    }
    return SPMV_OK;
}

// true when [y_begin, y_begin + y_count) is covered by this rank's parts (several of them: otherwise nothing overlaps)
static bool mg_can_pipeline(spmv_mg *g, int64_t y_begin, int64_t y_count)
{
    if (y_count <= 0 || g->parts.size() < 2) return false;
    std::vector<std::pair<int64_t, int64_t>> iv;
    for (const spmv_mg::Part &pt : g->parts) iv.push_back({pt.off, pt.off + pt.plan->N});
    std::sort(iv.begin(), iv.end());
    int64_t reach = y_begin;
    for (const auto &r : iv) {
        if (r.first > reach) break;
        reach = std::max(reach, r.second);
    }
    if (reach < y_begin + y_count) return false;
    while (g->part_done.size() < g->parts.size()) {
        cudaEvent_t e = nullptr;
        if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); return false; }
        g->part_done.push_back(e);
    }
    return true;
}

extern "C" {

int spmv_mg_run(spmv_mg_t *g, const float *d_x, void *stream, const float **d_y) { return mg_run_step(g, d_x, stream, d_y, false); }

int spmv_mg_status(spmv_mg_t *g)
{
    if (!g) return set_error(SPMV_ERR_ARG, "null group");
    unsigned s = 0;
    SPMV_CUDA(cudaMemcpy(&s, g->status, sizeof s, cudaMemcpyDeviceToHost));
    if (s) return set_error(SPMV_ERR_CUDA, "a peer rank did not arrive within %.0f s (spmv_mg join timed out)", kMgTimeoutNs * 1e-9);
    return SPMV_OK;
}

int spmv_mg_run_host(spmv_mg_t *g, const float *x, float *y, int64_t y_begin, int64_t y_count)
{
    if (!g) return set_error(SPMV_ERR_ARG, "null group");
    if ((!x && g->M > 0) || (!y && y_count > 0)) return set_error(SPMV_ERR_ARG, "null host vector");
    if (y_begin < 0 || y_count < 0 || y_begin + y_count > g->N) return set_error(SPMV_ERR_ARG, "y range outside [0, N)");
    if (g->M > 0) SPMV_CUDA(cudaMemcpyAsync(g->d_x, x, (size_t)g->M * sizeof(float), cudaMemcpyHostToDevice, g->stream));
    const float *dy = nullptr;
    const bool pipe = mg_can_pipeline(g, y_begin, y_count);
    int rc = mg_run_step(g, g->d_x, g->stream, &dy, pipe);
    if (!rc) rc = mg_copy_back(g, dy, y, y_begin, y_count, pipe);
    const cudaError_t e1 = cudaStreamSynchronize(g->stream), e2 = pipe ? cudaStreamSynchronize(g->copy_stream) : cudaSuccess;
    if (rc) return rc;
    if (e1 != cudaSuccess) return cuda_error(e1, "cudaStreamSynchronize");
    if (e2 != cudaSuccess) return cuda_error(e2, "cudaStreamSynchronize(copy stream)");
    return SPMV_OK;
}

} // extern "C"

// plans and groups are created on the calling thread's current device; a host program without the
// CUDA runtime headers (the drop-in launchers) selects it through this
extern "C" int spmv_set_device(int device)
{
    SPMV_CUDA(cudaSetDevice(device));
    return SPMV_OK;
}

// ---- several devices of ONE process (the reference's harness is a single process) -----------------
// Every device gets its own group handle (rank = position in `devices`); the blocks are plain
// cudaMalloc memory made mutually accessible with peer access, so the epilogues' stores to a
// peer's y and the flag words travel over NVLink exactly as in the one-process-per-GPU case.
extern "C" int spmv_mg_create_group(int64_t M, int64_t N_total, int n_dev, const int *devices, spmv_mg_t **out)
{
    if (!out || !devices || n_dev < 1 || n_dev > kMaxYDst) return set_error(SPMV_ERR_ARG, "1..%d devices required", kMaxYDst);
    for (int i = 0; i < n_dev; i++) out[i] = nullptr;
    int before = 0;
    cudaGetDevice(&before);
    int rc = SPMV_OK;
    for (int i = 0; i < n_dev && !rc; i++) {
        cudaError_t e = cudaSetDevice(devices[i]);
        if (e != cudaSuccess) { rc = cuda_error(e, "cudaSetDevice"); break; }
        for (int k = 0; k < n_dev; k++) {
            if (k == i) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, devices[i], devices[k]);
            if (!can) { rc = set_error(SPMV_ERR_UNSUPPORTED, "device %d cannot access device %d", devices[i], devices[k]); break; }
            e = cudaDeviceEnablePeerAccess(devices[k], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { rc = cuda_error(e, "cudaDeviceEnablePeerAccess"); break; }
            cudaGetLastError();
        }
        if (!rc) rc = spmv_mg_create(M, N_total, i, n_dev, nullptr, &out[i]);
    }
    if (!rc) {
        void *blocks[kMaxYDst] = {};
        for (int i = 0; i < n_dev; i++) blocks[i] = out[i]->block;
        for (int i = 0; i < n_dev && !rc; i++) rc = spmv_mg_connect_ptrs(out[i], blocks, nullptr);
    }
    if (rc) for (int i = 0; i < n_dev; i++) { if (out[i]) { cudaSetDevice(out[i]->device); spmv_mg_destroy(out[i]); out[i] = nullptr; } }
    cudaSetDevice(before);
    cudaGetLastError();
    return rc;
}

// One call on all devices of a group from one host thread: x to every device, every device's plans
// and arrival kernel (all asynchronous, so a device spinning for its peers never blocks the host
// from launching them), then every device's own columns of y back to the host.
extern "C" int spmv_mg_group_run_host(spmv_mg_t *const *groups, int n_dev, const float *x, float *y)
{
    if (!groups || n_dev < 1) return set_error(SPMV_ERR_ARG, "null argument");
    int before = 0;
    cudaGetDevice(&before);
    int rc = SPMV_OK;
    std::vector<const float *> dy((size_t)n_dev, nullptr);
    for (int i = 0; i < n_dev && !rc; i++) {
        spmv_mg *g = groups[i];
        if (!g) { rc = set_error(SPMV_ERR_ARG, "null group"); break; }
        cudaSetDevice(g->device);
        if (g->M > 0 && cudaMemcpyAsync(g->d_x, x, (size_t)g->M * sizeof(float), cudaMemcpyHostToDevice, g->stream) != cudaSuccess)
            rc = cuda_error(cudaGetLastError(), "cudaMemcpyAsync(x)");
    }
    std::vector<char> pipe((size_t)n_dev, 0);
    for (int i = 0; i < n_dev && !rc; i++) {
        spmv_mg *g = groups[i];
        cudaSetDevice(g->device);
        // every part's columns are wanted, so the pipeline condition is just "several parts" (events are made here)
        pipe[(size_t)i] = g->parts.size() > 1 && mg_can_pipeline(g, g->parts[0].off, 1);
        rc = mg_run_step(g, g->d_x, g->stream, &dy[(size_t)i], pipe[(size_t)i] != 0);
    }
    for (int i = 0; i < n_dev && !rc; i++) {                // own columns only: the union over the devices is all of y
        spmv_mg *g = groups[i];
        cudaSetDevice(g->device);
        for (size_t k = 0; k < g->parts.size() && !rc; k++) {
            const spmv_mg::Part &pt = g->parts[k];
            if (pt.plan->N <= 0) continue;
            cudaStream_t cs = g->stream;
            if (pipe[(size_t)i]) {
                cs = g->copy_stream;
                if (cudaStreamWaitEvent(cs, g->part_done[k], 0) != cudaSuccess) { rc = cuda_error(cudaGetLastError(), "cudaStreamWaitEvent"); break; }
            }
            if (cudaMemcpyAsync(y + pt.off, dy[(size_t)i] + pt.off, (size_t)pt.plan->N * sizeof(float), cudaMemcpyDeviceToHost, cs) != cudaSuccess)
                rc = cuda_error(cudaGetLastError(), "cudaMemcpyAsync(y)");
        }
    }
    for (int i = 0; i < n_dev; i++) {
        if (!groups[i]) continue;
        cudaSetDevice(groups[i]->device);
        cudaError_t e = cudaStreamSynchronize(groups[i]->stream);
        if (e == cudaSuccess && pipe[(size_t)i]) e = cudaStreamSynchronize(groups[i]->copy_stream);
        if (e != cudaSuccess && !rc) rc = cuda_error(e, "cudaStreamSynchronize");
        if (!rc) rc = spmv_mg_status(groups[i]);
    }
    cudaSetDevice(before);
    return rc;
}
