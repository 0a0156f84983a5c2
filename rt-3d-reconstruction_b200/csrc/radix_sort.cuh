// radix_sort.cuh -- internal API of the hand-written LSD radix sort (radix_sort.cu).
//
// The reference sorts voxel rows with torch.unique(dim=0) -- a thrust merge sort with a
// lexicographic row comparator (2cam/vision_pipeline_utils.py:43; its dominant geometric cost,
// SURVEY.md 8a V1).  Here rows become packed integer keys and are sorted by a one-sweep LSD radix
// sort: one histogram read of the keys up front, then one kernel per 8-bit digit that ranks a
// tile with warp match_any, resolves its global digit offsets by decoupled look-back and scatters
// through shared memory.  Element counts and the number of significant key bits may live in
// device memory, so the number of *executed* passes is decided on the GPU.
#pragma once

#include "rt3d_common.cuh"

namespace rt3d {

constexpr int kRsThreads = 256;
constexpr int kRsItems = 16;
constexpr int kRsTile = kRsThreads * kRsItems;  // 4096 keys per tile
constexpr int kRsMaxPasses = 8;

struct SortArgs {
    unsigned long long *keys[2];  // ping-pong buffers; input in keys[0]
    unsigned *vals[2];            // optional 32-bit payload (both nullptr for keys only)
    long long n;                  // key count, or capacity when n_dev != nullptr
    const int *n_dev;
    int total_bits;               // significant key bits (host value) ...
    const int *total_bits_dev;    // ... or device-resident (total_bits is then the upper bound)
    void *state;                  // sort_state_bytes(n, ...) bytes, zeroed by the caller
    int *status;                  // Ctrl::status
    bool hist_done;               // the producer of the keys already filled the digit histograms
};

inline int sort_max_passes(int total_bits) { return (total_bits + 7) / 8; }

// layout of `state`: [tickets: 16 ints][hist: 8*256 uint][tile digit states: passes * tiles * 256 uint]
inline size_t sort_state_bytes(long long n_capacity, int total_bits_bound) {
    const size_t tiles = (size_t)((n_capacity + kRsTile - 1) / kRsTile) + 1;
    return align_up(16 * sizeof(int)) + align_up(kRsMaxPasses * 256 * sizeof(unsigned)) +
           align_up((size_t)sort_max_passes(total_bits_bound) * tiles * 256 * sizeof(unsigned));
}
inline unsigned *sort_hist_ptr(void *state) { return (unsigned *)((char *)state + align_up(16 * sizeof(int))); }

// After the sort the keys (and payload) live in keys[p & 1], p = ceil(total_bits / 8) executed passes.
int launch_sort(const SortArgs &a, cudaStream_t stream);

#ifdef __CUDACC__
// Digit histograms of one key into a block-private [8][256] shared array.  Lanes holding the same
// digit are aggregated with match_any first: voxel keys of neighbouring pixels share their high
// digits, which would otherwise serialise on one shared-memory bank.
__device__ __forceinline__ void rs_hist_add(unsigned *s_hist, unsigned long long key, bool valid, int npass) {
    for (int p = 0; p < npass; ++p) {
        const unsigned d = valid ? (unsigned)((key >> (8 * p)) & 255u) : 256u + (unsigned)lane_id();
        const unsigned peers = __match_any_sync(kFullMask, d);
        if (valid && (__ffs(peers) - 1) == lane_id()) atomicAdd(&s_hist[p * 256 + d], (unsigned)__popc(peers));
    }
}
#endif

}  // namespace rt3d
