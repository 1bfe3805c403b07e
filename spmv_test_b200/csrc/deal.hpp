// deal.hpp — the order of the entries inside one chunk of 32 groups, shared by the host packers
// (pack_host.cpp) and the device packers (pack_dev.cu) so both produce the same bytes.
//
// Inside a WSP list, or a row segment of the panel formats, the order of the entries is free (any
// fixed order is deterministic), so it is chosen for the kernels' shared-memory accesses: within a
// chunk of 32 groups the 32 lanes handle element e (0..3) of their group in one instruction — a
// gather of x[row] in wsp, a read-modify-write of acc[column] in awsp/tcsr — and the entries are
// dealt so that those 32 indices fall into distinct banks (index mod 32) as far as the chunk allows.
#pragma once

#if defined(__CUDACC__)
#define SPMV_HD __host__ __device__ __forceinline__
#else
#define SPMV_HD inline
#endif

namespace spmv {

// One chunk of `lanes` groups (4*lanes entries, lanes <= 32).  key(k) is entry k's index; put(slot, k)
// stores entry k at position `slot` of the chunk.  The entries are bucketed by bank (ascending
// position inside a bucket), the banks are visited largest first (stable), and every entry goes to
// the element slot e that holds the fewest entries of its bank so far (ties: the emptier slot, then
// the lower e): position 4*fill[e] + e.
template <class Key, class Put>
SPMV_HD void deal_chunk(int lanes, Key key, Put put)
{
    const int count = 4 * lanes;
    unsigned char cnt[32], start[32], pos[32], order[32], sorted[128];
    for (int b = 0; b < 32; b++) cnt[b] = 0;
    for (int k = 0; k < count; k++) cnt[key(k) & 31u]++;
    int s = 0;
    for (int b = 0; b < 32; b++) { start[b] = (unsigned char)s; pos[b] = (unsigned char)s; s += cnt[b]; }
    for (int k = 0; k < count; k++) sorted[pos[key(k) & 31u]++] = (unsigned char)k;
    for (int b = 0; b < 32; b++) {                        // stable insertion sort, descending size
        int j = b;
        while (j > 0 && cnt[order[j - 1]] < cnt[b]) { order[j] = order[j - 1]; j--; }
        order[j] = (unsigned char)b;
    }
    int fill[4] = {0, 0, 0, 0};
    for (int t = 0; t < 32; t++) {
        const int b = order[t];
        const int nb = cnt[b];
        if (nb == 0) break;                               // the rest are empty too
        int mine[4] = {0, 0, 0, 0};
        for (int i = 0; i < nb; i++) {
            const int k = sorted[start[b] + i];
            int best = -1, bm = 0, bf = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int e = 0; e < 4; e++) {
                if (fill[e] >= lanes) continue;
                if (best < 0 || mine[e] < bm || (mine[e] == bm && fill[e] < bf)) { best = e; bm = mine[e]; bf = fill[e]; }
            }
            put(4 * bf + best, k);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int e = 0; e < 4; e++)
                if (e == best) { fill[e]++; mine[e]++; }
        }
    }
}

} // namespace spmv
