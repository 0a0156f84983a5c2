// select.cuh -- single-pass *ordered* stream compaction skeleton.
//
// Every "keep the rows that pass a test, in their original order" step of the reference (torch
// boolean indexing / torch.nonzero, each a cub select + host sync there) runs through this one
// kernel template: a persistent grid takes tiles by ticket, each warp evaluates a contiguous
// chunk (4 consecutive items per lane per row -> 128-bit loads), ranks survivors with warp
// shuffles, and one warp resolves the tile's global offset by decoupled look-back.  One read of
// the input, one write of the output, no host round trip; the survivor count lands in device
// memory.
//
// An Op provides
//     using Payload = ...;                                   // what a survivor carries to store()
//     __device__ unsigned eval4(long long i0, long long n, Payload (&p)[4]) const;
//         // items i0..i0+3 (i0 % 4 == 0); returns a 4-bit keep mask, fills p[j] for kept items
//     __device__ void store(long long pos, const Payload &p) const;   // pos = rank among survivors
#pragma once

#include "rt3d_common.cuh"

namespace rt3d {

constexpr int kSelThreads = 256;

struct SelectCtl {
    long long n;               // item count (host value), or capacity when n_dev != nullptr
    const int *n_dev;          // optional device-resident item count
    int *out_count;            // survivor count
    unsigned long long *state; // look-back words, one per tile, zeroed by the caller
    int *ticket;               // zeroed by the caller
    int *status;               // Ctrl::status
};

template <int ROWS>
constexpr int select_tile_items() { return kSelThreads * 4 * ROWS; }

template <class Op, int ROWS>
__global__ void __launch_bounds__(kSelThreads) select_kernel(const Op op, const SelectCtl c) {
    constexpr int TILE = kSelThreads * 4 * ROWS;
    constexpr int CHUNK = 32 * 4 * ROWS;  // items per warp per tile
    __shared__ int s_tile;
    __shared__ long long s_base;
    __shared__ int s_wtot[8];
    __shared__ int s_woff[8];

    const long long n = c.n_dev ? (long long)*c.n_dev : c.n;
    const int num_tiles = (int)((n + TILE - 1) / TILE);
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    if (num_tiles == 0) {
        if (blockIdx.x == 0 && threadIdx.x == 0) *c.out_count = 0;
        return;
    }
    for (;;) {
        if (threadIdx.x == 0) s_tile = atomicAdd(c.ticket, 1);
        __syncthreads();
        const int tile = s_tile;
        if (tile >= num_tiles) break;

        const long long chunk0 = (long long)tile * TILE + (long long)warp * CHUNK;
        typename Op::Payload pay[ROWS][4];
        unsigned keep[ROWS];
        int row_excl[ROWS];
        int warp_total = 0;
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            const long long i0 = chunk0 + r * 128 + lane * 4;
            keep[r] = (i0 < n) ? op.eval4(i0, n, pay[r]) : 0u;
        }
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            const int cnt = __popc(keep[r]);
            const int incl = warp_incl_scan(cnt);
            row_excl[r] = warp_total + incl - cnt;
            warp_total += __shfl_sync(kFullMask, incl, 31);
        }
        if (lane == 0) s_wtot[warp] = warp_total;
        __syncthreads();
        if (warp == 0) {
            const int wt = lane < 8 ? s_wtot[lane] : 0;
            const int wincl = warp_incl_scan(wt);
            const int tile_total = __shfl_sync(kFullMask, wincl, 7);
            const long long excl = lookback_exclusive(c.state, tile, tile_total, c.status);
            if (lane < 8) s_woff[lane] = wincl - wt;
            if (lane == 0) {
                s_base = excl;
                if (tile == num_tiles - 1) *c.out_count = (int)(excl + tile_total);
            }
        }
        __syncthreads();
        const long long base = s_base + s_woff[warp];
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            long long pos = base + row_excl[r];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if ((keep[r] >> j) & 1u) op.store(pos++, pay[r][j]);
        }
        __syncthreads();  // s_tile / s_base are rewritten by the next iteration
    }
}

// Host launcher.  `state` must hold select_state_words(n) zeroed words; `ticket` a zeroed int.
inline size_t select_state_words(long long n_capacity) {
    // sized for the smallest tile (ROWS = 1)
    return (size_t)((n_capacity + select_tile_items<1>() - 1) / select_tile_items<1>()) + 1;
}

template <class Op>
cudaError_t launch_select(const Op &op, const SelectCtl &c, cudaStream_t stream) {
    // small inputs: small tiles so that one image still spreads over the 148 SMs;
    // large (batched) inputs: 4 rows per lane to amortise the look-back.
    if (c.n <= 0) {
        // a device-side n may still be zero; capacity zero means nothing to do at all
        if (c.n_dev == nullptr) return cudaMemsetAsync(c.out_count, 0, sizeof(int), stream);
    }
    const long long n = c.n > 0 ? c.n : 1;
    if (n <= (long long)select_tile_items<1>() * kNumSMs * 8) {
        const long long tiles = (n + select_tile_items<1>() - 1) / select_tile_items<1>();
        select_kernel<Op, 1><<<persistent_blocks(tiles, 8), kSelThreads, 0, stream>>>(op, c);
    } else if (n <= (long long)select_tile_items<2>() * kNumSMs * 32) {
        const long long tiles = (n + select_tile_items<2>() - 1) / select_tile_items<2>();
        select_kernel<Op, 2><<<persistent_blocks(tiles, 8), kSelThreads, 0, stream>>>(op, c);
    } else {
        const long long tiles = (n + select_tile_items<4>() - 1) / select_tile_items<4>();
        select_kernel<Op, 4><<<persistent_blocks(tiles, 6), kSelThreads, 0, stream>>>(op, c);
    }
    return cudaGetLastError();
}

}  // namespace rt3d
