// subtract.cuh -- workspace subtraction (subtract_point_clouds_gpu, 2cam/vision_pipeline_utils.py:168-187)
// for one frame or a batch of frames; shared by k4_fuse_subtract.cu (implementation) and frame_batch.cu.
//
// The reference materialises the dense M x N fp32 cdist matrix and takes a row-min.  Here each
// frame's object cloud is binned into a three-level hierarchy of TIGHT boxes: 16x16x16 leaf cells
// (points radix-sorted by cell, a dozen points each) grouped 2x2x2 into 8x8x8 fine cells and again
// into 4x4x4 super cells.  A workspace point tests the frame's object box, the super cells, the
// fine cells of ambiguous super cells, the leaves of ambiguous fine cells, and finally scans the
// few points of the leaves that straddle the thr-sphere:
//     box farther than thr  -> skip,    box entirely within thr -> removed,
// with the final per-point decision made exactly (fp64 on the fp32 coordinates) when fp32 cannot
// separate the distance from the threshold.  The predicate is  keep <=> min_j |w - o_j| > thr.
#pragma once

#include <float.h>

#include "rt3d_common.cuh"

namespace rt3d {

constexpr int kSubFine = 8;                                // fine cells per axis (max)
constexpr int kSubFineCells = kSubFine * kSubFine * kSubFine;  // 512
constexpr int kSubSuper = 4;
constexpr int kSubSuperCells = kSubSuper * kSubSuper * kSubSuper;  // 64
constexpr int kSubLeaves = 8;                                     // 2x2x2 leaf cells per fine cell
constexpr int kSubLeafCells = kSubFineCells * kSubLeaves;          // 4096
constexpr int kSubCellBits = 12;                                   // sort key = frame | fine cell (9) | octant (3)

struct SubCell {  // 32 bytes: tight box of the points in the cell + their run in the cell-sorted order
    float lo[3], hi[3];
    int start, end;
};

struct SubRep {   // one actual point of a (non-empty) fine or super cell: within thr of it => within thr of the cloud
    float x, y, z, pad;
};

struct SubGrid {  // one frame's object cloud
    float lo[3], hi[3];     // tight box of the whole cloud
    float inv_cell[3];      // fine cells per metre, per axis
    int start, n;           // range in the object array
    int pad;
};

struct SubArgs {
    int F;                      // frames
    // objects: frame f owns rows [obj_off[f * off_stride], obj_off[(f + 1) * off_stride]) of `obj`
    const float *obj;
    const int *obj_off;
    int off_stride;
    long long obj_cap;          // capacity (grid sizing); actual total = obj_off[F * off_stride]
    // workspace: rows [ws_off[f * C], ws_off[(f + 1) * C]) belong to frame f
    const float *ws;
    const int *ws_off;
    int C;
    long long ws_cap;
    const int *ws_n_dev;        // total workspace rows (device)
    double t2;                  // squared radius: a row is removed iff some object point has d^2 <= t2 (< t2 if strict)
    int strict;
    // outputs
    unsigned char *keep;        // ws_cap flags (1 = kept)
    int *tile_cnt = nullptr;    // optional, zeroed: kept rows per tile of (1 << tile_shift) rows, for the compaction that follows
    int tile_shift = 0;
    // scratch
    unsigned long long *keys[2];
    unsigned *vals[2];
    void *sort_state;           // sort_state_bytes(obj_cap, kSubCellBits + frame bits), zeroed
    SubGrid *grid;              // F
    SubCell *fine;              // F * 512
    SubCell *leaf;              // F * 4096: leaf o of fine cell c at [c * 8 + o]
    SubCell *super;             // F * 64
    SubRep *fine_rep;           // F * 512
    SubRep *super_rep;          // F * 64
    float4 *obj_sorted;         // obj_cap points in cell-sorted order, one 16-byte load each (w unused)
    int *status;
    int *ticket;                // zeroed: 32-row chunks of the workspace are handed out dynamically
};

// number of key bits the object sort needs for F frames
inline int subtract_key_bits(int F) {
    int fb = 1;
    while ((1 << fb) < F) ++fb;
    return kSubCellBits + fb;
}

// fills a.keep[i] for every workspace row; frames without objects keep everything
int launch_subtract(const SubArgs &a, cudaStream_t stream);

#ifdef __CUDACC__
__device__ __forceinline__ void box_dist2(const float *lo, const float *hi, float x, float y, float z, float &dmin2, float &dmax2) {
    const float q[3] = {x, y, z};
    dmin2 = 0.f; dmax2 = 0.f;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float outside = fmaxf(fmaxf(lo[a] - q[a], q[a] - hi[a]), 0.f);
        const float far_ = fmaxf(q[a] - lo[a], hi[a] - q[a]);
        dmin2 = fmaf(outside, outside, dmin2);
        dmax2 = fmaf(far_, far_, dmax2);
    }
}

__device__ __forceinline__ float rep_dist2(const SubRep r, float x, float y, float z) {
    const float dx = x - r.x, dy = y - r.y, dz = z - r.z;
    return dx * dx + dy * dy + dz * dz;
}

// exact decision for one object point: fp32 first, fp64 on the fp32 coordinates when fp32 cannot separate the
// distance from the threshold
__device__ __forceinline__ bool point_within(float x, float y, float z, const float4 o, float t2f, double t2, bool strict) {
    const float dx = x - o.x, dy = y - o.y, dz = z - o.z;
    const float d2f = dx * dx + dy * dy + dz * dz;
    if (d2f > t2f * (1.0f + 4e-6f)) return false;
    if (d2f < t2f * (1.0f - 4e-6f)) return true;
    const double ex = (double)x - (double)o.x, ey = (double)y - (double)o.y, ez = (double)z - (double)o.z;
    const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey)), __dmul_rn(ez, ez));
    return strict ? d2 < t2 : d2 <= t2;
}

// points start, start + step, ... below end (RT3D_SUB_PREFETCH: the next point is in flight while the current one is tested)
#ifndef RT3D_SUB_PREFETCH
#define RT3D_SUB_PREFETCH 1
#endif
__device__ __forceinline__ bool scan_points(float x, float y, float z, const float4 *__restrict__ obj_sorted, int start, int end,
                                            float t2f, double t2, bool strict, int step = 1) {
#if RT3D_SUB_PREFETCH
    if (start >= end) return false;
    const float4 *p = obj_sorted + start;
    float4 cur = *p;
    for (int j = start + step; j < end; j += step) {
        p += step;
        const float4 nxt = *p;
        if (point_within(x, y, z, cur, t2f, t2, strict)) return true;
        cur = nxt;
    }
    return point_within(x, y, z, cur, t2f, t2, strict);
#else
    for (int j = start; j < end; j += step)
        if (point_within(x, y, z, obj_sorted[j], t2f, t2, strict)) return true;
    return false;
#endif
}

#ifndef RT3D_SUB_FINE8_MAX
#define RT3D_SUB_FINE8_MAX 32
#endif
constexpr int kSubFine8Max = RT3D_SUB_FINE8_MAX;   // longest fine-cell list worked with eight lanes per item
// Per-warp work lists of the row traversal.  The deeper levels of the hierarchy are where the rows of a warp part
// ways (a leaf's points were being scanned with two lanes active on average), so a row does not descend itself:
// it files (row, fine cell) and (row, leaf run) items, and the warp then works the lists with one item per lane.
constexpr int kSubQueueFine = 128, kSubQueueRuns = 192;
struct SubQueues {
    unsigned long long runs[kSubQueueRuns];   // start (32) | count (16) | row (8)
    unsigned fine[kSubQueueFine];             // fine cell << 5 | row
    int n_fine, n_runs;
    unsigned near;                            // bit r: row r has been found near
};

// the leaves of one fine cell for one row; true = near.  Straddling leaves are filed in q (or scanned when q is
// full / absent).
__device__ __forceinline__ bool scan_leaves(float x, float y, float z, int row, const SubCell *__restrict__ leaf, int cell,
                                            const float4 *__restrict__ obj_sorted, float t2f, double t2, bool strict, SubQueues *q) {
    const float t2_hi = t2f * (1.0f + 1e-5f), t2_lo = t2f * (1.0f - 1e-5f);
#pragma unroll 1
    for (int o = 0; o < kSubLeaves; ++o) {
        const SubCell lc = leaf[cell * kSubLeaves + o];
        if (lc.start >= lc.end) continue;
        float dmin2, dmax2;
        box_dist2(lc.lo, lc.hi, x, y, z, dmin2, dmax2);
        if (dmax2 < t2_lo) return true;
        if (dmin2 > t2_hi) continue;
        if (q) {
            const int slot = atomicAdd(&q->n_runs, 1);
            if (slot < kSubQueueRuns) {
                q->runs[slot] = ((unsigned long long)(unsigned)lc.start << 32) | ((unsigned long long)(lc.end - lc.start) << 8) | (unsigned)row;
                continue;
            }
        }
        if (scan_points(x, y, z, obj_sorted, lc.start, lc.end, t2f, t2, strict)) return true;
    }
    return false;
}

// true iff some object point of this frame lies within thr of (x,y,z) -- or, with work lists (q), false plus the
// items whose outcome is still open.  One row per lane.
__device__ __forceinline__ bool near_object(const SubGrid &g, const SubCell *__restrict__ super, const SubCell *__restrict__ fine,
                                            const SubCell *__restrict__ leaf, const SubRep *__restrict__ super_rep,
                                            const SubRep *__restrict__ fine_rep, const float4 *__restrict__ obj_sorted, float t2f,
                                            double t2, bool strict, float x, float y, float z,
                                            unsigned long long candidates = ~0ull, const unsigned *fine_candidates = nullptr,
                                            SubQueues *q = nullptr, int row = 0) {
    // fp32 box distances carry ~1e-6 relative error: decide only with a safety margin
    const float t2_hi = t2f * (1.0f + 1e-5f), t2_lo = t2f * (1.0f - 1e-5f);
    float dmin2, dmax2;
    box_dist2(g.lo, g.hi, x, y, z, dmin2, dmax2);
    if (dmin2 > t2_hi) return false;
    if (dmax2 < t2_lo) return true;
    // every super cell, accept first: a row within thr of the cell's representative point is done -- most rows
    // near an object are decided here
    const float t2_acc = t2f * (1.0f - 4e-6f);
    // (`candidates`: the super cells the caller has not already excluded for this row -- warp-uniform;
    // `fine_candidates`: likewise one byte of fine-cell bits for each of them, in bit order, four per word)
    unsigned long long ambiguous = 0ull;
    const unsigned long long all_candidates = candidates;
    for (; candidates; candidates &= candidates - 1) {
        const int s = __ffsll((long long)candidates) - 1;
        const SubCell sc = super[s];
        if (sc.start >= sc.end) continue;                 // empty
        if (rep_dist2(super_rep[s], x, y, z) < t2_acc) return true;
        box_dist2(sc.lo, sc.hi, x, y, z, dmin2, dmax2);
        if (dmin2 <= t2_hi) ambiguous |= 1ull << s;
    }
    for (; ambiguous; ambiguous &= ambiguous - 1) {
        const int s = __ffsll((long long)ambiguous) - 1;
        const int sx = s >> 4, sy = (s >> 2) & 3, sz = s & 3;
        unsigned cells = 0xffu;
        if (fine_candidates) {
            const int idx = __popcll(all_candidates & ((1ull << s) - 1ull));
            cells = (fine_candidates[idx >> 2] >> (8 * (idx & 3))) & 0xffu;
        }
#pragma unroll 1
        for (; cells; cells &= cells - 1) {
            const int c8 = __ffs((int)cells) - 1;
            const int cell = (((2 * sx + (c8 >> 2)) * kSubFine) + (2 * sy + ((c8 >> 1) & 1))) * kSubFine + (2 * sz + (c8 & 1));
            const SubCell fc = fine[cell];
            if (fc.start >= fc.end) continue;
            if (rep_dist2(fine_rep[cell], x, y, z) < t2_acc) return true;
            box_dist2(fc.lo, fc.hi, x, y, z, dmin2, dmax2);
            if (dmin2 > t2_hi) continue;
            // a straddling fine cell: its leaves go on the list
            if (q) {
                const int slot = atomicAdd(&q->n_fine, 1);
                if (slot < kSubQueueFine) { q->fine[slot] = ((unsigned)cell << 5) | (unsigned)row; continue; }
            }
            if (scan_leaves(x, y, z, row, leaf, cell, obj_sorted, t2f, t2, strict, q)) return true;
        }
    }
    return false;
}

// The warp works its lists: one filed item per lane.  x, y, z: the lane's own row (rows are addressed by lane).
__device__ __forceinline__ void drain_queues(SubQueues *q, const SubCell *__restrict__ leaf, const float4 *__restrict__ obj_sorted,
                                             float t2f, double t2, bool strict, float x, float y, float z) {
    const int lane = lane_id();
    __syncwarp();
    // fine cells.  A short list (the common case at 5 mm / thr 0.3: a dozen items) is worked with eight lanes per item, one
    // leaf each -- one lane per item would leave most of the warp idle.  A long list (2 mm voxels, many objects, thr 0.06:
    // the list is usually full) keeps one lane per item walking its eight leaves: every lane is busy anyway, and the walk
    // stops at the first leaf that settles the row (measured at BASELINE configs[2]: 2.81 vs 3.98 ms for the whole stage).
    const int nf = min(q->n_fine, kSubQueueFine);
    if (nf <= kSubFine8Max) {
        int n_runs = q->n_runs;   // runs filed so far (warp-uniform: the lists are private to the warp and its lanes are
                                  // converged here, so appending takes a vote, not shared-memory atomics on one counter)
        const float t2_hi = t2f * (1.0f + 1e-5f), t2_lo = t2f * (1.0f - 1e-5f);
        const int o = lane & 7, g8 = lane >> 3;
        for (int base = 0; base < nf; base += 4) {
            const int it = base + g8;
            const unsigned item = it < nf ? q->fine[it] : 0u;
            const int row = item & 31;
            const float rx = __shfl_sync(kFullMask, x, row), ry = __shfl_sync(kFullMask, y, row), rz = __shfl_sync(kFullMask, z, row);
            bool want = false;
            int ps = 0, pe = 0;
            if (it < nf && !((*(volatile unsigned *)&q->near >> row) & 1u)) {
                const SubCell lc = leaf[(int)(item >> 5) * kSubLeaves + o];
                if (lc.start < lc.end) {
                    float dmin2, dmax2;
                    box_dist2(lc.lo, lc.hi, rx, ry, rz, dmin2, dmax2);
                    if (dmax2 < t2_lo) atomicOr(&q->near, 1u << row);
                    else if (dmin2 <= t2_hi) { want = true; ps = lc.start; pe = lc.end; }
                }
            }
            const unsigned wm = __ballot_sync(kFullMask, want);
            if (wm == 0u) continue;   // warp-uniform
            const int slot = n_runs + __popc(wm & lanemask_lt());
            n_runs += __popc(wm);
            if (want) {
                if (slot < kSubQueueRuns) q->runs[slot] = ((unsigned long long)(unsigned)ps << 32) | ((unsigned long long)(pe - ps) << 8) | (unsigned)row;
                else if (scan_points(rx, ry, rz, obj_sorted, ps, pe, t2f, t2, strict)) atomicOr(&q->near, 1u << row);
            }
        }
        if (lane == 0) q->n_runs = n_runs;
    } else {
        for (int base = 0; base < nf; base += 32) {
            const int it = base + lane;
            const unsigned item = it < nf ? q->fine[it] : 0u;
            const int row = item & 31;
            const float rx = __shfl_sync(kFullMask, x, row), ry = __shfl_sync(kFullMask, y, row), rz = __shfl_sync(kFullMask, z, row);
            if (it < nf && !((*(volatile unsigned *)&q->near >> row) & 1u))
                if (scan_leaves(rx, ry, rz, row, leaf, (int)(item >> 5), obj_sorted, t2f, t2, strict, q)) atomicOr(&q->near, 1u << row);
        }
    }
    __syncwarp();
    const int nr = min(q->n_runs, kSubQueueRuns);
    // leaf runs: four lanes per item, each taking every fourth point (a run holds a dozen points; with one lane per
    // run the warp waited for its longest run with a quarter of its lanes active)
    const int sub = lane & 3, grp = lane >> 2;
    for (int base = 0; base < nr; base += 8) {
        const int it = base + grp;
        const unsigned long long item = it < nr ? q->runs[it] : 0ull;
        const int row = (int)(item & 31ull), start = (int)(item >> 32), cnt = (int)((item >> 8) & 0xffffffull);
        const float rx = __shfl_sync(kFullMask, x, row), ry = __shfl_sync(kFullMask, y, row), rz = __shfl_sync(kFullMask, z, row);
        if (it < nr && !((*(volatile unsigned *)&q->near >> row) & 1u))
            if (scan_points(rx, ry, rz, obj_sorted, start + sub, start + cnt, t2f, t2, strict, 4)) atomicOr(&q->near, 1u << row);
    }
    __syncwarp();
}
#endif  // __CUDACC__

}  // namespace rt3d
