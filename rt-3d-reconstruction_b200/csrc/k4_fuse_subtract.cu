// k4_fuse_subtract.cu -- kernel family 4: centroids for the fusion matcher and workspace subtraction.
//
//  calculate_centroid (2cam/vision_pipeline_utils.py:74-75): np.mean(axis=0) = sequential row sum in
//      the array's dtype, then / N.  One warp, three lanes carry the x / y / z chains.
//  S1  subtract_point_clouds_gpu (:168-187): the reference materialises the dense M x N fp32 cdist
//      matrix (4.9 GB at config 2) and takes a row-min.  Here the object cloud is binned into at most
//      512 coarse cells with tight per-cell bounding boxes; each workspace point culls cells by
//      box distance (skip if surely farther than thr, accept if surely nearer) and only scans the
//      ambiguous ones, deciding near-threshold pairs exactly in fp64.  Traffic 12*M + 12*N in,
//      12*M' out instead of 4*M*N.
#include <float.h>
#include <limits.h>

#include "radix_sort.cuh"
#include "select.cuh"
#include "subtract.cuh"

namespace rt3d {

// ------------------------------------------------------------------------------- centroid
template <class T>
__global__ void centroid_kernel(const T *__restrict__ xyz, long long n, T *out3) {
    const int a = threadIdx.x;
    if (a >= 3) return;
    T s = (T)0;
    long long i = 0;
    for (; i + 8 <= n; i += 8) {  // 8 independent loads in flight, one dependent add chain
        T v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = xyz[3 * (i + u) + a];
#pragma unroll
        for (int u = 0; u < 8; ++u) s = sizeof(T) == 4 ? (T)__fadd_rn((float)s, (float)v[u]) : (T)__dadd_rn((double)s, (double)v[u]);
    }
    for (; i < n; ++i) {
        const T v = xyz[3 * i + a];
        s = sizeof(T) == 4 ? (T)__fadd_rn((float)s, (float)v) : (T)__dadd_rn((double)s, (double)v);
    }
    out3[a] = sizeof(T) == 4 ? (T)__fdiv_rn((float)s, (float)n) : (T)__ddiv_rn((double)s, (double)n);
}

// ------------------------------------------------------------------------------- subtraction
constexpr int kSubThreads = 256;

// one block per frame: tight box of the frame's object cloud -> fine-cell geometry
__global__ void __launch_bounds__(128) sub_prepare_kernel(const SubArgs a) {
    __shared__ float s_lo[3][4], s_hi[3][4];
    const int f = blockIdx.x;
    const int beg = a.obj_off[f * a.off_stride], end = a.obj_off[(f + 1) * a.off_stride];
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i = beg + threadIdx.x; i < end; i += 128) {
#pragma unroll
        for (int k = 0; k < 3; ++k) { const float v = a.obj[3 * (long long)i + k]; lo[k] = fminf(lo[k], v); hi[k] = fmaxf(hi[k], v); }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[k] = fminf(lo[k], __shfl_xor_sync(kFullMask, lo[k], o));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(kFullMask, hi[k], o));
        }
        if (lane_id() == 0) { s_lo[k][threadIdx.x >> 5] = lo[k]; s_hi[k][threadIdx.x >> 5] = hi[k]; }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        SubGrid g;
        for (int k = 0; k < 3; ++k) {
            float l = fminf(fminf(s_lo[k][0], s_lo[k][1]), fminf(s_lo[k][2], s_lo[k][3]));
            float h = fmaxf(fmaxf(s_hi[k][0], s_hi[k][1]), fmaxf(s_hi[k][2], s_hi[k][3]));
            if (!(l <= h)) { l = 0.f; h = 0.f; }
            g.lo[k] = l; g.hi[k] = h;
            const float ext = h - l;
            g.inv_cell[k] = ext > 0.f ? (float)kSubFine / ext : 0.f;   // 8 cells across the extent (1 if flat)
        }
        g.start = beg; g.n = end - beg; g.pad = 0;
        a.grid[f] = g;
    }
}

__device__ __forceinline__ int frame_of_row(const int *__restrict__ off, int stride, int F, int i) {
    int lo = 0, hi = F;  // largest f with off[f * stride] <= i
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (off[mid * stride] <= i) lo = mid; else hi = mid; }
    return lo;
}

__global__ void __launch_bounds__(kSubThreads) sub_cellkey_kernel(const SubArgs a, int npass, unsigned *hist) {
    __shared__ unsigned s_hist[kRsMaxPasses * 256];
    const int n = a.obj_off[a.F * a.off_stride];
    for (int i = threadIdx.x; i < kRsMaxPasses * 256; i += kSubThreads) s_hist[i] = 0;
    __syncthreads();
    const long long stride = (long long)gridDim.x * kSubThreads;
    for (long long base = (long long)blockIdx.x * kSubThreads; base < n; base += stride) {
        const long long i = base + threadIdx.x;
        const bool valid = i < n;
        unsigned long long key = 0;
        if (valid) {
            const int f = frame_of_row(a.obj_off, a.off_stride, a.F, (int)i);
            const SubGrid g = a.grid[f];
            int c[3];   // leaf cell (16 per axis); inv_cell counts fine cells, a leaf is half a fine cell
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                c[k] = (int)floorf((a.obj[3 * i + k] - g.lo[k]) * (2.0f * g.inv_cell[k]));
                c[k] = c[k] < 0 ? 0 : (c[k] >= 2 * kSubFine ? 2 * kSubFine - 1 : c[k]);
            }
            const int finec = ((c[0] >> 1) * kSubFine + (c[1] >> 1)) * kSubFine + (c[2] >> 1);
            const int oct = ((c[0] & 1) << 2) | ((c[1] & 1) << 1) | (c[2] & 1);
            key = ((unsigned long long)f << kSubCellBits) | (unsigned long long)(finec * kSubLeaves + oct);
            a.keys[0][i] = key;
            a.vals[0][i] = (unsigned)i;
        }
        rs_hist_add(s_hist, key, valid, npass);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < npass * 256; i += kSubThreads) {
        const unsigned c = s_hist[i];
        if (c) atomicAdd(&hist[i], c);
    }
}

// order-preserving float <-> int (so that warp integer min / max reductions order floats)
__device__ __forceinline__ int float_ordered(float f) { const int i = __float_as_int(f); return i ^ ((i >> 31) & 0x7fffffff); }
__device__ __forceinline__ float ordered_float(int i) { return __int_as_float(i ^ ((i >> 31) & 0x7fffffff)); }

// one warp per (frame, fine cell): its run in the sorted pairs, the points gathered in sorted order, and the
// tight boxes of its 8 leaves (points of a leaf are contiguous: the octant is the low 3 key bits) and of itself
__global__ void __launch_bounds__(kSubThreads) sub_cells_kernel(const SubArgs a, const unsigned long long *__restrict__ keys,
                                                                const unsigned *__restrict__ vals) {
    const int w = (blockIdx.x * kSubThreads + threadIdx.x) >> 5;
    const int f = w / kSubFineCells, cell = w - f * kSubFineCells;
    if (f >= a.F) return;
    const SubGrid g = a.grid[f];
    const int lane = lane_id();
    const unsigned long long key = (((unsigned long long)f << kSubCellBits) | ((unsigned long long)cell << 3));
    // the frame's points occupy [g.start, g.start + g.n) before and after the sort (frame id = high key bits);
    // lanes 0..8 locate the first key >= key + lane: the starts of the 8 leaves and the end of the cell
    long long l = g.start, h = (long long)g.start + g.n;
    const unsigned long long want = key + (unsigned long long)(lane < 9 ? lane : 8);
    while (l < h) { const long long m = (l + h) >> 1; if (keys[m] < want) l = m + 1; else h = m; }
    const long long lo_i = __shfl_sync(kFullMask, l, 0), hi_i = __shfl_sync(kFullMask, l, 8);
    int blo[3] = {INT_MAX, INT_MAX, INT_MAX}, bhi[3] = {INT_MIN, INT_MIN, INT_MIN};   // lane o < 8: box of leaf o (ordered ints)
    for (long long j0 = lo_i; j0 < hi_i; j0 += 32) {
        const long long j = j0 + lane;
        const bool valid = j < hi_i;
        int q[3] = {0, 0, 0}, oct = -1;
        if (valid) {
            const long long i = vals[j];
            oct = (int)(keys[j] & 7ull);
            const float vx = a.obj[3 * i], vy = a.obj[3 * i + 1], vz = a.obj[3 * i + 2];
            a.obj_sorted[j] = make_float4(vx, vy, vz, 0.f);
            q[0] = float_ordered(vx); q[1] = float_ordered(vy); q[2] = float_ordered(vz);
        }
#pragma unroll
        for (int o = 0; o < kSubLeaves; ++o) {
            const unsigned m = __ballot_sync(kFullMask, oct == o);
            if (m == 0) continue;   // warp-uniform
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int mn = __reduce_min_sync(kFullMask, oct == o ? q[k] : INT_MAX), mx = __reduce_max_sync(kFullMask, oct == o ? q[k] : INT_MIN);
                if (lane == o) { blo[k] = min(blo[k], mn); bhi[k] = max(bhi[k], mx); }
            }
        }
    }
    // leaves, then the fine cell = union of its leaves
    const long long my_s = l, my_e = __shfl_down_sync(kFullMask, l, 1);
    if (lane < kSubLeaves) {
        SubCell c;
        for (int k = 0; k < 3; ++k) { c.lo[k] = ordered_float(blo[k]); c.hi[k] = ordered_float(bhi[k]); }
        c.start = (int)my_s; c.end = (int)my_e;
        a.leaf[((size_t)f * kSubFineCells + cell) * kSubLeaves + lane] = c;
    }
    int ulo[3], uhi[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { ulo[k] = __reduce_min_sync(kFullMask, blo[k]); uhi[k] = __reduce_max_sync(kFullMask, bhi[k]); }
    if (lane == 0) {
        SubCell c;
        for (int k = 0; k < 3; ++k) { c.lo[k] = ordered_float(ulo[k]); c.hi[k] = ordered_float(uhi[k]); }
        c.start = (int)lo_i; c.end = (int)hi_i;
        a.fine[(size_t)f * kSubFineCells + cell] = c;
        SubRep r = {0.f, 0.f, 0.f, 0.f};
        if (lo_i < hi_i) { const long long i = vals[lo_i]; r.x = a.obj[3 * i]; r.y = a.obj[3 * i + 1]; r.z = a.obj[3 * i + 2]; }
        a.fine_rep[(size_t)f * kSubFineCells + cell] = r;
    }
}

// one thread per (frame, super cell): union of its 2x2x2 fine boxes
__global__ void sub_super_kernel(const SubArgs a) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int f = t / kSubSuperCells, s = t - f * kSubSuperCells;
    if (f >= a.F) return;
    const int sx = s >> 4, sy = (s >> 2) & 3, sz = s & 3;
    SubCell u;
    for (int k = 0; k < 3; ++k) { u.lo[k] = FLT_MAX; u.hi[k] = -FLT_MAX; }
    int count = 0;
    for (int c8 = 0; c8 < 8; ++c8) {
        const int cell = (((2 * sx + (c8 >> 2)) * kSubFine) + (2 * sy + ((c8 >> 1) & 1))) * kSubFine + (2 * sz + (c8 & 1));
        const SubCell fc = a.fine[(size_t)f * kSubFineCells + cell];
        if (fc.start >= fc.end) continue;
        count += fc.end - fc.start;
        for (int k = 0; k < 3; ++k) { u.lo[k] = fminf(u.lo[k], fc.lo[k]); u.hi[k] = fmaxf(u.hi[k], fc.hi[k]); }
    }
    u.start = 0; u.end = count;
    a.super[(size_t)f * kSubSuperCells + s] = u;
    // representative: the fine-cell representative nearest the centre of the super box
    SubRep best = {0.f, 0.f, 0.f, 0.f};
    float bd = FLT_MAX;
    for (int c8 = 0; c8 < 8 && count > 0; ++c8) {
        const int cell = (((2 * sx + (c8 >> 2)) * kSubFine) + (2 * sy + ((c8 >> 1) & 1))) * kSubFine + (2 * sz + (c8 & 1));
        const SubCell fc = a.fine[(size_t)f * kSubFineCells + cell];
        if (fc.start >= fc.end) continue;
        const SubRep r = a.fine_rep[(size_t)f * kSubFineCells + cell];
        const float d = rep_dist2(r, 0.5f * (u.lo[0] + u.hi[0]), 0.5f * (u.lo[1] + u.hi[1]), 0.5f * (u.lo[2] + u.hi[2]));
        if (d < bd) { bd = d; best = r; }
    }
    a.super_rep[(size_t)f * kSubSuperCells + s] = best;
}

// squared distance from a point to the farthest point of a box
__device__ __forceinline__ float box_point_far2(const float *alo, const float *ahi, const SubRep r) {
    const float fx = fmaxf(r.x - alo[0], ahi[0] - r.x), fy = fmaxf(r.y - alo[1], ahi[1] - r.y), fz = fmaxf(r.z - alo[2], ahi[2] - r.z);
    return fmaf(fx, fx, fmaf(fy, fy, fz * fz));
}

// squared distance bounds between two axis-aligned boxes: nearest and farthest pair of points
__device__ __forceinline__ void box_box_dist2(const float *alo, const float *ahi, const float *blo, const float *bhi, float &dmin2, float &dmax2) {
    dmin2 = 0.f; dmax2 = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float gap = fmaxf(fmaxf(blo[k] - ahi[k], alo[k] - bhi[k]), 0.f);
        const float far_ = fmaxf(ahi[k] - blo[k], bhi[k] - alo[k]);
        dmin2 = fmaf(gap, gap, dmin2);
        dmax2 = fmaf(far_, far_, dmax2);
    }
}

// 32 query points of ONE frame (one per lane, spatially close), decided TOGETHER where possible.  With A = their
// bounding box,
//   - every object cell farther than thr from A            -> none is near,
//   - some non-empty cell entirely within thr of all of A  -> all are near,
// tested against the frame's object box, then the 64 super cells (two per lane), then the 512 fine cells.  Only
// groups straddling the thr-shell around the object fall through to the per-point traversal of the three-level
// hierarchy (near_object), restricted to the cells the group tests left open.  All box bounds are decided with a
// safety margin; the exact fp64 predicate stays per point.  `fine_amb` (kSubSuperCells / 4 words) and `q` are the
// warp's shared-memory scratch.
__device__ __forceinline__ bool warp_near_object(const SubArgs &a, int f0, const SubGrid &g, float x, float y, float z, bool live,
                                                 float t2f, double t2, bool strict, unsigned *fine_amb, SubQueues *q, int lane) {
    const float t2_hi = t2f * (1.0f + 1e-5f), t2_lo = t2f * (1.0f - 1e-5f);
    int decided = g.n > 0 ? 0 : 1;   // 1: none near, 2: all near
    float alo[3], ahi[3];
    if (!decided) {
        alo[0] = ordered_float(__reduce_min_sync(kFullMask, float_ordered(x))); ahi[0] = ordered_float(__reduce_max_sync(kFullMask, float_ordered(x)));
        alo[1] = ordered_float(__reduce_min_sync(kFullMask, float_ordered(y))); ahi[1] = ordered_float(__reduce_max_sync(kFullMask, float_ordered(y)));
        alo[2] = ordered_float(__reduce_min_sync(kFullMask, float_ordered(z))); ahi[2] = ordered_float(__reduce_max_sync(kFullMask, float_ordered(z)));
        float dmin2, dmax2;
        box_box_dist2(alo, ahi, g.lo, g.hi, dmin2, dmax2);
        if (dmin2 > t2_hi) decided = 1;
        else if (dmax2 < t2_lo) decided = 2;
    }
    unsigned long long super_amb = ~0ull;   // super cells within thr of A
    if (!decided) {   // super cells: lane tests cells lane and lane + 32
        const SubCell *sup = a.super + (size_t)f0 * kSubSuperCells;
        const SubRep *srep = a.super_rep + (size_t)f0 * kSubSuperCells;
        bool acc = false;
        unsigned am[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const SubCell sc = sup[lane + 32 * h];
            bool amb = false;
            if (sc.start < sc.end) {
                float dmin2, dmax2;
                box_box_dist2(alo, ahi, sc.lo, sc.hi, dmin2, dmax2);
                acc = acc || box_point_far2(alo, ahi, srep[lane + 32 * h]) < t2_lo;
                amb = dmin2 <= t2_hi;
            }
            am[h] = __ballot_sync(kFullMask, amb);
        }
        super_amb = (unsigned long long)am[0] | ((unsigned long long)am[1] << 32);
        if (__any_sync(kFullMask, acc)) decided = 2;
        else if (super_amb == 0ull) decided = 1;
    }
    if (!decided) {   // fine cells
        const SubCell *fin = a.fine + (size_t)f0 * kSubFineCells;
        const SubRep *frep = a.fine_rep + (size_t)f0 * kSubFineCells;
        // the 8 fine cells of every super cell that is still ambiguous, four super cells per round
        bool acc = false, any_amb = false;
        int round = 0;
        for (unsigned long long m = super_amb; m; ++round) {
            unsigned long long mine = m;
            for (int qq = lane >> 3; qq > 0 && mine; --qq) mine &= mine - 1;
            bool amb = false;
            if (mine) {
                const int s = __ffsll((long long)mine) - 1, c8 = lane & 7;
                const int cell = (((2 * (s >> 4) + (c8 >> 2)) * kSubFine) + (2 * ((s >> 2) & 3) + ((c8 >> 1) & 1))) * kSubFine + (2 * (s & 3) + (c8 & 1));
                const SubCell fc = fin[cell];
                if (fc.start < fc.end) {
                    float dmin2, dmax2;
                    box_box_dist2(alo, ahi, fc.lo, fc.hi, dmin2, dmax2);
                    acc = acc || box_point_far2(alo, ahi, frep[cell]) < t2_lo;
                    amb = dmin2 <= t2_hi;
                }
            }
            const unsigned bits = __ballot_sync(kFullMask, amb);   // byte q = fine cells of the round's q-th super cell
            any_amb = any_amb || bits != 0u;
            if (lane == 0) fine_amb[round] = bits;
#pragma unroll
            for (int qq = 0; qq < 4; ++qq) m &= m - 1;
        }
        __syncwarp();
        if (__any_sync(kFullMask, acc)) decided = 2;
        else if (!any_amb) decided = 1;
    }
    if (decided) return decided == 2;
    if (lane == 0) { q->n_fine = 0; q->n_runs = 0; q->near = 0u; }
    __syncwarp();
    bool near = false;
    if (live)
        near = near_object(g, a.super + (size_t)f0 * kSubSuperCells, a.fine + (size_t)f0 * kSubFineCells,
                           a.leaf + (size_t)f0 * kSubLeafCells, a.super_rep + (size_t)f0 * kSubSuperCells,
                           a.fine_rep + (size_t)f0 * kSubFineCells, a.obj_sorted, t2f, t2, strict, x, y, z, super_amb, fine_amb, q, lane);
    if (near) atomicOr(&q->near, 1u << lane);
    drain_queues(q, a.leaf + (size_t)f0 * kSubLeafCells, a.obj_sorted, t2f, t2, strict, x, y, z);
    near = (q->near >> lane) & 1u;
    __syncwarp();   // the lists are reused by the warp's next group
    return near;
}

// image of row r -- the largest m with off[m] <= r -- found by the whole warp, 32 probes per round
__device__ __forceinline__ int image_of_row_warp(const int *__restrict__ off, int nimg, int r, int lane) {
    int lo = 0, span = nimg;   // the answer lies in [lo, lo + span)
    while (span > 1) {
        const int step = (span + 31) >> 5;
        const int m = lo + (lane + 1) * step;
        const int k = __popc(__ballot_sync(kFullMask, m < lo + span && off[m] <= r));
        lo += k * step;
        span = min(step, span - k * step);
    }
    return lo;
}

// One thread per workspace row, a warp per 32 consecutive rows.  Rows are sorted lexicographically, so a warp's
// rows form a short run of neighbouring voxels that warp_near_object mostly decides together.
#ifndef RT3D_SUB_BLOCKS
#define RT3D_SUB_BLOCKS 6   // blocks per SM of the flag kernel (tuning experiments override it)
#endif
__global__ void __launch_bounds__(kSubThreads, RT3D_SUB_BLOCKS) sub_flag_kernel(const SubArgs a, float t2f) {
    // per warp: the fine cells within thr of its box, one byte per ambiguous super cell
    __shared__ unsigned s_fine_amb[kSubThreads / 32][kSubSuperCells / 4];
    __shared__ SubQueues s_queues[kSubThreads / 32];
    const int n = *a.ws_n_dev;
    const int nimg = a.F * a.C;
    const int lane = lane_id();
    // 32-row chunks by atomic ticket (a chunk on the thr-shell costs 50x one far from it); next ticket in flight
    int tk = 0;
    if (lane == 0) tk = atomicAdd(a.ticket, 1);
    for (;;) {
        const long long base = 32ll * __shfl_sync(kFullMask, tk, 0);
        if (base >= n) break;
        if (lane == 0) tk = atomicAdd(a.ticket, 1);
        const long long i = base + lane;
        const bool live = i < n;
        const long long ic = live ? i : (long long)n - 1;
        const float x = a.ws[3 * ic], y = a.ws[3 * ic + 1], z = a.ws[3 * ic + 2];
        // rows are frame-major: the image of the first row, and whether the last row is still inside it
        const int last = (int)min(base + 31, (long long)n - 1);
        const int img0 = image_of_row_warp(a.ws_off, nimg, (int)base, lane);
        const int f0 = img0 / a.C;
        bool near = false;
        if (last < a.ws_off[(f0 + 1) * a.C]) {
            near = warp_near_object(a, f0, a.grid[f0], x, y, z, live, t2f, a.t2, a.strict != 0, s_fine_amb[threadIdx.x >> 5], &s_queues[threadIdx.x >> 5], lane);
        } else if (live) {   // the warp straddles a frame boundary: every row with its own frame
            const int fr = frame_of_row(a.ws_off, 1, nimg, (int)i) / a.C;
            const SubGrid g = a.grid[fr];
            if (g.n > 0)
                near = near_object(g, a.super + (size_t)fr * kSubSuperCells, a.fine + (size_t)fr * kSubFineCells,
                                   a.leaf + (size_t)fr * kSubLeafCells, a.super_rep + (size_t)fr * kSubSuperCells,
                                   a.fine_rep + (size_t)fr * kSubFineCells, a.obj_sorted, t2f, a.t2, a.strict != 0, x, y, z);
        }
        if (live) a.keep[i] = near ? 0 : 1;
        if (a.tile_cnt != nullptr) {   // 32-row groups never straddle a tile (tiles are multiples of 32 rows)
            const int kept = __popc(__ballot_sync(kFullMask, live && !near));
            if (lane == 0 && kept) atomicAdd(&a.tile_cnt[base >> a.tile_shift], kept);
        }
    }
}

int launch_subtract(const SubArgs &a, cudaStream_t stream) {
    RT3D_REQUIRE(a.F >= 1 && a.F <= 65535 && a.C >= 1 && a.t2 >= 0.0, RT3D_EINVAL, "subtract: bad arguments");
    const int bits = subtract_key_bits(a.F);
    const int npass = sort_max_passes(bits);
    const int gobj = persistent_blocks((a.obj_cap + kSubThreads - 1) / kSubThreads, 8);
    sub_prepare_kernel<<<a.F, 128, 0, stream>>>(a);
    sub_cellkey_kernel<<<gobj, kSubThreads, 0, stream>>>(a, npass, sort_hist_ptr(a.sort_state));
    RT3D_LAUNCH_OK();
    SortArgs sa;
    sa.keys[0] = a.keys[0]; sa.keys[1] = a.keys[1]; sa.vals[0] = a.vals[0]; sa.vals[1] = a.vals[1];
    sa.n = a.obj_cap; sa.n_dev = a.obj_off + (size_t)a.F * a.off_stride; sa.total_bits = bits; sa.total_bits_dev = nullptr;
    sa.state = a.sort_state; sa.status = a.status; sa.hist_done = true;
    int rc = launch_sort(sa, stream);
    if (rc) return rc;
    sub_cells_kernel<<<(a.F * kSubFineCells * 32 + kSubThreads - 1) / kSubThreads, kSubThreads, 0, stream>>>(a, a.keys[npass & 1], a.vals[npass & 1]);
    sub_super_kernel<<<(a.F * kSubSuperCells + 127) / 128, 128, 0, stream>>>(a);
    sub_flag_kernel<<<persistent_blocks((a.ws_cap + kSubThreads - 1) / kSubThreads, 8), kSubThreads, 0, stream>>>(a, (float)a.t2);
    RT3D_LAUNCH_OK();
    return RT3D_OK;
}

// ordered compaction of the rows whose flag is set
struct KeepRowsOp {
    struct Payload { float x, y, z; };
    const float *ws;
    const unsigned char *keep;
    float *out_xyz;
    __device__ __forceinline__ unsigned eval4(long long i0, long long n, Payload (&p)[4]) const {
        unsigned k = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const long long i = i0 + j;
            if (i >= n) break;
            if (keep[i]) { p[j].x = ws[3 * i]; p[j].y = ws[3 * i + 1]; p[j].z = ws[3 * i + 2]; k |= 1u << j; }
        }
        return k;
    }
    __device__ __forceinline__ void store(long long pos, const Payload &p) const {
        if (!out_xyz) return;
        float *o = out_xyz + 3 * pos;
        o[0] = p.x; o[1] = p.y; o[2] = p.z;
    }
};

__global__ void sub_fill_offsets_kernel(int *obj_off, int n, int *ws_off, int m, int *ws_n) {
    obj_off[0] = 0; obj_off[1] = n; ws_off[0] = 0; ws_off[1] = m; *ws_n = m;
}

struct SubPlan {
    Ctrl *ctrl;
    void *sort_state;
    unsigned long long *sel_state;
    size_t zero_bytes;
    int *offsets;   // obj_off[2], ws_off[2], ws_n
    unsigned long long *keys[2];
    unsigned *vals[2];
    SubGrid *grid;
    SubCell *fine, *super, *leaf;
    SubRep *fine_rep, *super_rep;
    float4 *obj_sorted;
    unsigned char *keep;
    size_t total;
    bool ok;
};

static SubPlan plan_sub(void *scratch, size_t bytes, long long m, long long n) {
    Arena a(scratch, bytes);
    SubPlan p;
    p.ctrl = a.take<Ctrl>(1);
    p.sort_state = a.take<char>(sort_state_bytes(n, subtract_key_bits(1)));
    p.sel_state = a.take<unsigned long long>(select_state_words(m));
    p.zero_bytes = a.off;
    p.offsets = a.take<int>(8);
    p.keys[0] = a.take<unsigned long long>((size_t)n + 4);
    p.keys[1] = a.take<unsigned long long>((size_t)n + 4);
    p.vals[0] = a.take<unsigned>((size_t)n + 4);
    p.vals[1] = a.take<unsigned>((size_t)n + 4);
    p.grid = a.take<SubGrid>(1);
    p.fine = a.take<SubCell>(kSubFineCells);
    p.leaf = a.take<SubCell>(kSubLeafCells);
    p.super = a.take<SubCell>(kSubSuperCells);
    p.fine_rep = a.take<SubRep>(kSubFineCells);
    p.super_rep = a.take<SubRep>(kSubSuperCells);
    p.obj_sorted = a.take<float4>((size_t)n + 1);
    p.keep = a.take<unsigned char>((size_t)m + 4);
    p.total = a.off;
    p.ok = a.ok();
    return p;
}

size_t subtract_scratch_bytes(long long m, long long n) { return plan_sub(nullptr, 0, m, n).total; }

}  // namespace rt3d

using namespace rt3d;

extern "C" int rt3d_centroid(const void *xyz, int dtype, int64_t n, void *out3, rt3d_stream_t stream_) {
    RT3D_REQUIRE(xyz && out3 && n >= 0, RT3D_EINVAL, "rt3d_centroid: bad arguments");
    RT3D_REQUIRE(dtype == RT3D_F32 || dtype == RT3D_F64, RT3D_EINVAL, "dtype must be RT3D_F32 or RT3D_F64");
    if (dtype == RT3D_F32) centroid_kernel<float><<<1, 32, 0, (cudaStream_t)stream_>>>((const float *)xyz, n, (float *)out3);
    else centroid_kernel<double><<<1, 32, 0, (cudaStream_t)stream_>>>((const double *)xyz, n, (double *)out3);
    RT3D_LAUNCH_OK();
    return RT3D_OK;
}

static int subtract_impl(const float *ws, int64_t m, const float *obj, int64_t n, double t2, int strict, uint8_t *out_keep,
                         float *out_xyz, int32_t *out_count, void *scratch, size_t scratch_bytes, cudaStream_t stream);

extern "C" int rt3d_subtract(const float *ws, int64_t m, const float *obj, int64_t n, float thr, uint8_t *out_keep,
                             float *out_xyz, int32_t *out_count, void *scratch, size_t scratch_bytes,
                             rt3d_stream_t stream_) {
    RT3D_REQUIRE(thr >= 0.0f, RT3D_EINVAL, "distance threshold must be non-negative");
    return subtract_impl(ws, m, obj, n, (double)thr * (double)thr, 0, out_keep, out_xyz, out_count, scratch, scratch_bytes,
                         (cudaStream_t)stream_);
}

extern "C" int rt3d_subtract_radius(const float *ws, int64_t m, const float *obj, int64_t n, double radius, uint8_t *out_keep,
                                    float *out_xyz, int32_t *out_count, void *scratch, size_t scratch_bytes,
                                    rt3d_stream_t stream_) {
    RT3D_REQUIRE(radius >= 0.0, RT3D_EINVAL, "radius must be non-negative");
    return subtract_impl(ws, m, obj, n, radius * radius, 1, out_keep, out_xyz, out_count, scratch, scratch_bytes,
                         (cudaStream_t)stream_);
}

static int subtract_impl(const float *ws, int64_t m, const float *obj, int64_t n, double t2, int strict, uint8_t *out_keep,
                         float *out_xyz, int32_t *out_count, void *scratch, size_t scratch_bytes, cudaStream_t stream) {
    RT3D_REQUIRE((ws || m == 0) && obj && out_count && scratch, RT3D_EINVAL, "rt3d_subtract: bad arguments");
    RT3D_REQUIRE(m >= 0 && m < (1ll << 30) && n < (1ll << 30), RT3D_ERANGE, "point counts outside [0, 2^30)");
    RT3D_REQUIRE(n >= 1, RT3D_EINVAL, "rt3d_subtract: empty object cloud (the reference raises IndexError here)");
    SubPlan p = plan_sub(scratch, scratch_bytes, m, n);
    RT3D_REQUIRE(p.ok, RT3D_ESCRATCH, "scratch too small: need %zu bytes, have %zu", p.total, scratch_bytes);
    RT3D_CUDA_OK(cudaMemsetAsync(scratch, 0, p.zero_bytes, stream));
    if (m == 0) { RT3D_CUDA_OK(cudaMemsetAsync(out_count, 0, sizeof(int), stream)); return RT3D_OK; }
    sub_fill_offsets_kernel<<<1, 1, 0, stream>>>(p.offsets, (int)n, p.offsets + 2, (int)m, p.offsets + 4);
    SubArgs a;
    a.F = 1; a.obj = obj; a.obj_off = p.offsets; a.off_stride = 1; a.obj_cap = n;
    a.ws = ws; a.ws_off = p.offsets + 2; a.C = 1; a.ws_cap = m; a.ws_n_dev = p.offsets + 4; a.t2 = t2; a.strict = strict;
    a.keep = out_keep ? out_keep : p.keep;
    a.keys[0] = p.keys[0]; a.keys[1] = p.keys[1]; a.vals[0] = p.vals[0]; a.vals[1] = p.vals[1];
    a.sort_state = p.sort_state; a.grid = p.grid; a.fine = p.fine; a.leaf = p.leaf; a.super = p.super; a.fine_rep = p.fine_rep; a.super_rep = p.super_rep; a.ticket = &p.ctrl->ticket[1]; a.obj_sorted = p.obj_sorted;
    a.status = &p.ctrl->status;
    int rc = launch_subtract(a, stream);
    if (rc) return rc;
    KeepRowsOp op{ws, a.keep, out_xyz};
    SelectCtl c;
    c.n = m; c.n_dev = nullptr; c.out_count = out_count; c.state = p.sel_state; c.ticket = &p.ctrl->ticket[0];
    c.status = &p.ctrl->status;
    RT3D_CUDA_OK(launch_select(op, c, stream));
    return RT3D_OK;
}
