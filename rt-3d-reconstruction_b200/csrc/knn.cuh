// knn.cuh -- uniform-grid k-nearest-neighbour search shared by the stand-alone outlier filter
// (k3_sor.cu) and the batched frame pipeline (frame_batch.cu).
//
// Points are sorted by a cell key  prefix | cx | cy | cz  (z minor), so one (x, y) column of cells
// over a z range is ONE contiguous run of the sorted keys, found with two binary searches.  A warp
// serves one query: lanes locate the columns of the (2r+1)^2 neighbourhood in parallel, then
// stream the candidates 32 at a time into a warp-distributed sorted list of the k smallest
// squared distances (ballot to filter, shuffles to insert).  Distances are fp64 in nanoflann's
// operation order; the final sum runs ascending, as Open3D does.
#pragma once

#include <float.h>

#include "rt3d_common.cuh"

namespace rt3d {

struct GridGeom {
    double lo[3], h, inv_h;
    int dim[3], bits[3];
    unsigned long long prefix;  // high key bits shared by every point of this cloud (batched: the segment id)
    const int *cols = nullptr;  // optional: cols[x * dim[1] + y] = first sorted index of column (x, y) or later;
                                // dim[0] * dim[1] + 1 entries, the last one = end of the cloud
    __device__ __forceinline__ int cell(double p, int a) const {
        int c = (int)floor((p - lo[a]) * inv_h);
        return c < 0 ? 0 : (c >= dim[a] ? dim[a] - 1 : c);
    }
    __device__ __forceinline__ unsigned long long pack(int cx, int cy, int cz) const {
        return prefix | ((unsigned long long)cx << (bits[1] + bits[2])) | ((unsigned long long)cy << bits[2]) |
               (unsigned long long)cz;
    }
    __device__ __forceinline__ void unpack(unsigned long long k, int &cx, int &cy, int &cz) const {
        cz = (int)(k & ((1ull << bits[2]) - 1ull));
        cy = (int)((k >> bits[2]) & ((1ull << bits[1]) - 1ull));
        cx = (int)((k >> (bits[1] + bits[2])) & ((1ull << bits[0]) - 1ull));
    }
};

// The bitonic networks below (15 + 5 compare-exchange stages, inlined at every place a batch of candidates is offered) are
// kept as rolled loops: unrolled, the search kernels grow past 100 KB of code and the batched kernel spent 40 % of its
// warp time waiting for instruction fetches (ncu: stall_no_instruction 7.0 warps per issue).  RT3D_KNN_ROLL 0 unrolls them.
#ifndef RT3D_KNN_ROLL
#define RT3D_KNN_ROLL 1
#endif
#if RT3D_KNN_ROLL
#define RT3D_KNN_SORT_UNROLL _Pragma("unroll 1")
#else
#define RT3D_KNN_SORT_UNROLL _Pragma("unroll")
#endif
// ---- warp-distributed sorted list of the k smallest values (k <= 32 * SLOTS), ascending
template <int SLOTS>
struct WarpTopK {
    double v[SLOTS];
    double kth;   // current kk-th smallest value (DBL_MAX while fewer than kk are held); warp-uniform
    __device__ __forceinline__ void reset() {
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) v[s] = DBL_MAX;
        kth = DBL_MAX;
    }
    __device__ __forceinline__ double at(int p) const {  // p warp-uniform
        double r = 0.0;
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) {
            const double t = __shfl_sync(kFullMask, v[s], p & 31);
            if ((p >> 5) == s) r = t;
        }
        return r;
    }
    __device__ __forceinline__ void insert(double x) {  // x warp-uniform
        int pos = 0;
#pragma unroll
        for (int s = 0; s < SLOTS; ++s) pos += __popc(__ballot_sync(kFullMask, v[s] <= x));
        const int lane = lane_id();
#pragma unroll
        for (int s = SLOTS - 1; s >= 0; --s) {
            double up = __shfl_up_sync(kFullMask, v[s], 1);
            if (s > 0) {
                const double carry = __shfl_sync(kFullMask, v[s - 1], 31);
                if (lane == 0) up = carry;
            }
            const int gp = s * 32 + lane;
            if (gp == pos) v[s] = x;
            else if (gp > pos) v[s] = up;
        }
    }
    // compare-exchange with the lane at distance j: the lower lane of a pair keeps the minimum when `up`
    static __device__ __forceinline__ double cex(double x, int j, bool up) {
        const double other = __shfl_xor_sync(kFullMask, x, j);
        const bool lower = (lane_id() & j) == 0;
        return (lower == up) ? fmin(x, other) : fmax(x, other);
    }
    // offer one candidate per lane; kk = number of neighbours wanted
    __device__ __forceinline__ void offer(double d2, bool valid, int kk) {
        unsigned need = __ballot_sync(kFullMask, valid && d2 < kth);
        if (SLOTS == 1 && __popc(need) > 2) {
            // many newcomers: sort the batch (bitonic, 15 stages), lay it reversed against the held list, keep
            // the element-wise minima (a bitonic sequence holding the 32 smallest of the 64) and merge (5 stages).
            // Only VALUES are ranked, so equal distances need no tie-break.
            const int lane = lane_id();
            double x = (valid && d2 < kth) ? d2 : DBL_MAX;
            RT3D_KNN_SORT_UNROLL
            for (int k2 = 2; k2 <= 32; k2 <<= 1) {
                RT3D_KNN_SORT_UNROLL
                for (int j = k2 >> 1; j > 0; j >>= 1) x = cex(x, j, (lane & k2) == 0);
            }
            double m = fmin(v[0], __shfl_sync(kFullMask, x, 31 - lane));
            RT3D_KNN_SORT_UNROLL
            for (int j = 16; j > 0; j >>= 1) m = cex(m, j, true);
            v[0] = m;
            kth = at(kk - 1);
            return;
        }
        while (need) {
            const int src = __ffs(need) - 1;
            need &= need - 1;
            const double x = __shfl_sync(kFullMask, d2, src);
            if (x < kth) {
                insert(x);
                kth = at(kk - 1);
            }
        }
    }
};

__device__ __forceinline__ double dist2_nanoflann(double qx, double qy, double qz, const double *p) {
    const double dx = __dsub_rn(qx, p[0]), dy = __dsub_rn(qy, p[1]), dz = __dsub_rn(qz, p[2]);
    return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

__device__ __forceinline__ int lower_bound_u64(const unsigned long long *a, int lo, int hi, unsigned long long key) {
    while (lo < hi) {
        const int mid = (int)(((unsigned)lo + (unsigned)hi) >> 1);
        if (a[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// [s, e): the run of sorted keys in cells (x, y, zlo..zhi).  With a column table the search is confined to the
// column's own run (a handful of keys); without one it is two binary searches over the whole cloud.
__device__ __forceinline__ void column_run(const GridGeom &g, const unsigned long long *__restrict__ keys, int seg_lo, int seg_hi,
                                           int x, int y, int zlo, int zhi, int &s, int &e) {
    if (g.cols != nullptr) {
        const int c = x * g.dim[1] + y;
        const int cs = g.cols[c], ce = g.cols[c + 1];
        s = zlo <= 0 ? cs : lower_bound_u64(keys, cs, ce, g.pack(x, y, zlo));
        e = zhi >= g.dim[2] - 1 ? ce : lower_bound_u64(keys, s, ce, g.pack(x, y, zhi) + 1ull);
    } else {
        s = lower_bound_u64(keys, seg_lo, seg_hi, g.pack(x, y, zlo));
        e = lower_bound_u64(keys, s, seg_hi, g.pack(x, y, zhi) + 1ull);
    }
}

// Per-warp shared scratch of the fast path.
constexpr int kKnnFastCap = 304, kKnnBins = 64;
struct KnnWarpScratch {
    double d2[kKnnFastCap];     // candidate distances
    int idx[kKnnFastCap];       // candidate indices (the runs of the block's columns, concatenated)
    unsigned char bin[kKnnFastCap];   // histogram bin of each candidate (kKnnBins = beyond reach)
    double sel[32];             // the selected (then sorted) ones
    int hist[kKnnBins];
};

// Fast path for the common case: the (2r+1)^3 cell block around the query (r = 1, then 2, then 4) holds at most
// kKnnFastCap candidates, the kk-th neighbour lies inside its guaranteed reach r*h, and kk <= 32.  Select
// instead of insert:
//   1. one lane per (x, y) column of the block locates the column's z run (contiguous in the sorted keys) and
//      lists its indices; the warp then computes the exact fp64 distances 32 at a time into shared memory and
//      a 64-bin histogram of [0, reach^2];
//   2. the first bin B whose cumulative count reaches kk fixes the candidate set {bin <= B} -- a superset of
//      the kk smallest, because binning is monotone in the distance;
//   3. if that set has at most 32 members they are ranked exactly (counting sort on the fp64 values, one per
//      lane), square-rooted and summed ascending, as Open3D does.
// Returns 0 when `mean` is final; 1 when the r = 4 block holds too many candidates for the list (the caller's
// general search starts at r = 4); 3 when even the r = 4 block does not reach the kk-th neighbour (exhaustive
// scan); 2 for any other unmet precondition (general search from the start).
// `cache` (warp-uniform, optional): the candidate list left in ws->idx by the previous query of this warp.  Queries
// arrive in cell order, so the next one usually sits in the same cell and reuses the list (any r whose block
// reaches the kk-th neighbour gives the exact answer, so a query may start at its predecessor's r).
struct KnnCache {
    unsigned long long key;   // sorted key (cloud prefix | cell) the list was built for; ~0: none
    int r, total;
};

__device__ __forceinline__ int knn_mean_distance_fast(const GridGeom &g, const unsigned long long *__restrict__ keys,
                                                      const double *__restrict__ ps, int seg_lo, int seg_hi, double qx, double qy,
                                                      double qz, int cx, int cy, int cz, int kk, KnnWarpScratch *ws, double &mean,
                                                      KnnCache *cache = nullptr, unsigned long long qkey = 0ull, int r_start = 1) {
    const int lane = lane_id();
    if (kk > 32) return 2;
    const bool hit = cache != nullptr && cache->key == qkey;
    for (int r = hit ? max(cache->r, r_start) : r_start; r <= 4; r <<= 1) {
        const int side = 2 * r + 1, ncols = side * side;
        // one lane per (x, y) column of the block, 32 columns per round; the runs' indices are appended to the list
        int total = 0;
        bool overflow = false;
        const bool reuse = hit && cache->key == qkey && cache->r == r;
        if (reuse) total = cache->total;
        for (int cb = 0; cb < ncols && !overflow && !reuse; cb += 32) {
            const int ci = cb + lane;
            int s1 = 0, e1 = 0;
            if (ci < ncols) {
                const int x = cx + ci / side - r, y = cy + ci % side - r;
                if (x >= 0 && x < g.dim[0] && y >= 0 && y < g.dim[1]) {
                    const int zlo = max(cz - r, 0), zhi = min(cz + r, g.dim[2] - 1);
                    column_run(g, keys, seg_lo, seg_hi, x, y, zlo, zhi, s1, e1);
                }
            }
            const int len = e1 - s1;
            const int incl = warp_incl_scan(len);
            const int round_total = __shfl_sync(kFullMask, incl, 31);
            if (total + round_total > kKnnFastCap) { overflow = true; break; }
            int at_ = total + incl - len;
            for (int t = s1; t < e1; ++t) ws->idx[at_++] = t;
            total += round_total;
        }
        if (overflow) {                        // too many candidates for the list: general search
            if (cache) cache->key = ~0ull;
            return r == 4 ? 1 : 2;
        }
        if (cache && !reuse) { cache->key = qkey; cache->r = r; cache->total = total; }
        if (total < kk) continue;
        const double reach = (double)r * g.h * (1.0 - 1e-9), reach2 = reach * reach;
        const double bin_scale = (double)kKnnBins / reach2;
        for (int b = lane; b < kKnnBins; b += 32) ws->hist[b] = 0;
        __syncwarp();
        for (int base = 0; base < total; base += 32) {
            const int t = base + lane;
            if (t < total) {
                const double d2 = dist2_nanoflann(qx, qy, qz, ps + 3 * (size_t)ws->idx[t]);
                ws->d2[t] = d2;
                int b = kKnnBins;
                if (d2 <= reach2) { b = min((int)(d2 * bin_scale), kKnnBins - 1); atomicAdd(&ws->hist[b], 1); }
                ws->bin[t] = (unsigned char)b;
            }
        }
        __syncwarp();
        // first bin whose cumulative count reaches kk
        const int h0 = ws->hist[lane], h1 = ws->hist[lane + 32];
        const int c0 = warp_incl_scan(h0);
        const int tot0 = __shfl_sync(kFullMask, c0, 31);
        const int c1 = tot0 + warp_incl_scan(h1);
        const unsigned m0 = __ballot_sync(kFullMask, c0 >= kk), m1 = __ballot_sync(kFullMask, c1 >= kk);
        if ((m0 | m1) == 0) continue;                        // fewer than kk neighbours within reach: widen
        const int B = m0 ? (__ffs(m0) - 1) : 32 + (__ffs(m1) - 1);
        const int count = m0 ? __shfl_sync(kFullMask, c0, B) : __shfl_sync(kFullMask, c1, B - 32);
        if (count > 32) return 2;                            // too many ties around the kk-th distance
        // gather {bin <= B}
        ws->sel[lane] = DBL_MAX;
        __syncwarp();
        int filled = 0;
        for (int base = 0; base < total; base += 32) {
            const int t = base + lane;
            const bool take = t < total && (int)ws->bin[t] <= B;
            const unsigned tm = __ballot_sync(kFullMask, take);
            if (take) ws->sel[filled + __popc(tm & lanemask_lt())] = ws->d2[t];
            filled += __popc(tm);
        }
        __syncwarp();
        // exact ranks (ties by position), one value per lane
        const double mine = ws->sel[lane];
        int rank = 0;
#pragma unroll 8
        for (int j = 0; j < 32; ++j) {
            const double o = ws->sel[j];
            rank += (o < mine || (o == mine && j < lane)) ? 1 : 0;
        }
        __syncwarp();
        ws->sel[rank] = sqrt(mine);      // sqrt(DBL_MAX) for the padding: never read below
        __syncwarp();
        double sum = 0.0;
        for (int p = 0; p < kk; ++p) sum = __dadd_rn(sum, ws->sel[p]);   // ascending, from 0.0, like Open3D
        mean = __ddiv_rn(sum, (double)kk);
        __syncwarp();
        return 0;
    }
    return 3;   // not even the r = 4 block reaches the kk-th neighbour: only the exhaustive scan can
}

// Mean distance from query (qx,qy,qz) -- which lives in cell (cx,cy,cz) of `g` -- to its kk nearest
// neighbours (itself included) among the points whose sorted keys occupy [seg_lo, seg_hi).
// `keys` / `ps` are the sorted keys and the points in the same order.  Called by a full warp; every
// lane returns the same value.
// `s_list`: kKnnListCap ints of shared memory private to the calling warp (or nullptr).
// `r_start` (1, 2 or 4): the caller knows that smaller blocks cannot hold the kk neighbours (any block that reaches the
// kk-th neighbour gives the exact answer, so a larger start only costs candidates).
constexpr int kKnnListCap = 512;
template <int SLOTS>
__device__ __forceinline__ double knn_mean_distance(const GridGeom &g, const unsigned long long *__restrict__ keys,
                                                    const double *__restrict__ ps, int seg_lo, int seg_hi,
                                                    double qx, double qy, double qz, int cx, int cy, int cz, int kk, int *s_list,
                                                    KnnWarpScratch *ws = nullptr, KnnCache *cache = nullptr, unsigned long long qkey = 0ull,
                                                    int r_start = 1) {
    const int lane = lane_id();
    int r_first = r_start;
    if (SLOTS == 1 && ws != nullptr) {
        double mean;
        const int rc = knn_mean_distance_fast(g, keys, ps, seg_lo, seg_hi, qx, qy, qz, cx, cy, cz, kk, ws, mean, cache, qkey, r_start);
        if (rc == 0) return mean;
        if (rc == 1) r_first = 4;   // nothing nearer than r = 4 can settle it: scan the r = 4 block in one go
        if (rc == 3) r_first = 8;   // straight to the exhaustive scan
    }
    WarpTopK<SLOTS> top;
    top.reset();
    bool done = false;
    int r_prev = -1;  // cells within Chebyshev distance r_prev of the query cell are already in the list
    for (int r = r_first; r <= 4 && !done; r <<= 1) {
        const int side = 2 * r + 1, ncols = side * side;
        for (int cb = 0; cb < ncols; cb += 32) {
            // each lane locates the not-yet-scanned part of one column (x, y): one z run for a new column,
            // two (below / above the scanned block) for a column of the previous ring.  A z run of a
            // column is a contiguous run of the sorted keys.
            const int ci = cb + lane;
            int s1 = 0, e1 = 0, s2 = 0, e2 = 0;
            if (ci < ncols) {
                const int dx = ci / side - r, dy = ci % side - r;
                const int x = cx + dx, y = cy + dy;
                if (x >= 0 && x < g.dim[0] && y >= 0 && y < g.dim[1]) {
                    const bool inner = r_prev >= 0 && abs(dx) <= r_prev && abs(dy) <= r_prev;
                    int zlo = max(cz - r, 0), zhi = min(inner ? cz - r_prev - 1 : cz + r, g.dim[2] - 1);
                    if (zlo <= zhi) column_run(g, keys, seg_lo, seg_hi, x, y, zlo, zhi, s1, e1);
                    if (inner) {
                        zlo = max(cz + r_prev + 1, 0); zhi = min(cz + r, g.dim[2] - 1);
                        if (zlo <= zhi) column_run(g, keys, seg_lo, seg_hi, x, y, zlo, zhi, s2, e2);
                    }
                }
            }
            // pack the runs of all 32 columns into full 32-candidate batches through the warp's index list
            const int len = (e1 - s1) + (e2 - s2);
            const int incl = warp_incl_scan(len);
            const int total = __shfl_sync(kFullMask, incl, 31);
            if (s_list != nullptr && total <= kKnnListCap) {
                int at_ = incl - len;
                for (int t = s1; t < e1; ++t) s_list[at_++] = t;
                for (int t = s2; t < e2; ++t) s_list[at_++] = t;
                __syncwarp();
                for (int base = 0; base < total; base += 32) {
                    const bool valid = base + lane < total;
                    const double d2 = valid ? dist2_nanoflann(qx, qy, qz, ps + 3 * (size_t)s_list[base + lane]) : DBL_MAX;
                    top.offer(d2, valid, kk);
                }
                __syncwarp();
            } else {
                const unsigned nonempty = __ballot_sync(kFullMask, len > 0);
                for (unsigned m = nonempty; m; m &= m - 1) {
                    const int c = __ffs(m) - 1;
                    RT3D_KNN_SORT_UNROLL
                    for (int part = 0; part < 2; ++part) {
                        const int cs = __shfl_sync(kFullMask, part ? s2 : s1, c), ce = __shfl_sync(kFullMask, part ? e2 : e1, c);
                        for (int base = cs; base < ce; base += 32) {
                            const int t = base + lane;
                            const bool valid = t < ce;
                            const double d2 = valid ? dist2_nanoflann(qx, qy, qz, ps + 3 * (size_t)t) : DBL_MAX;
                            top.offer(d2, valid, kk);
                        }
                    }
                }
            }
        }
        r_prev = r;
        // every point outside the scanned box is farther than r*h (minus rounding slack)
        const double reach = (double)r * g.h * (1.0 - 1e-9);
        const bool covers_all = cx - r <= 0 && cx + r >= g.dim[0] - 1 && cy - r <= 0 && cy + r >= g.dim[1] - 1 &&
                                cz - r <= 0 && cz + r >= g.dim[2] - 1;
        done = covers_all || (top.kth < DBL_MAX && top.kth <= reach * reach);
    }
    if (!done) {  // isolated point: exhaustive scan of its cloud, in two passes
        // pass 1: each lane's smallest distance.  The kk-th smallest of the 32 lane minima bounds the kk-th smallest
        // distance overall from above (kk <= 32 values are at or below it), so pass 2 offers only the candidates
        // under that bound to the sorted list -- a few dozen instead of the whole cloud.
        top.reset();
        double bound = DBL_MAX;
        if (SLOTS == 1 && seg_hi - seg_lo >= 32 * 8) {
            double mn = DBL_MAX;
            for (int t = seg_lo + lane; t < seg_hi; t += 32) mn = fmin(mn, dist2_nanoflann(qx, qy, qz, ps + 3 * (size_t)t));
            RT3D_KNN_SORT_UNROLL
            for (int k2 = 2; k2 <= 32; k2 <<= 1) {
                RT3D_KNN_SORT_UNROLL
                for (int j = k2 >> 1; j > 0; j >>= 1) mn = WarpTopK<SLOTS>::cex(mn, j, (lane & k2) == 0);
            }
            bound = __shfl_sync(kFullMask, mn, kk - 1);   // ascending by lane after the sort
        }
        for (int base = seg_lo; base < seg_hi; base += 32) {
            const int t = base + lane;
            const double d2 = t < seg_hi ? dist2_nanoflann(qx, qy, qz, ps + 3 * (size_t)t) : DBL_MAX;
            const bool valid = t < seg_hi && d2 <= bound;
            if (__ballot_sync(kFullMask, valid) == 0) continue;
            top.offer(d2, valid, kk);
        }
    }
    // Open3D: sqrt each, accumulate ascending from 0.0, divide by the neighbour count
    double sum = 0.0;
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
        const double root = sqrt(top.v[s]);  // one square root per lane, in parallel
        const int cnt = min(32, kk - 32 * s);
        for (int p = 0; p < cnt; ++p) sum = __dadd_rn(sum, __shfl_sync(kFullMask, root, p));
    }
    return __ddiv_rn(sum, (double)kk);
}

}  // namespace rt3d
