// frame_batch.cu -- the whole per-frame geometric path for a BATCH of multi-camera frames, device
// resident from depth maps to subtracted workspace, 40-47 launches per batch and no host round trip.
//
// One call does, for F frames x C cameras (2cam/2cams.py:165-208, geometry only):
//   A  project   one pass over every depth map and its instance masks: back-project, rigid
//                transform (process_masks :292-304 / process_point_clouds :229-254), then
//                  - workspace: crop (:78-84), voxel index, one bit in a per-image two-level
//                    occupancy bitmap of the crop box (the voxel set of :40-44, built by red.or);
//                  - objects: packed (segment | voxel) keys appended to one stream.
//   B  extract   ordered scan of the occupancy bitmaps -> each image's sorted unique lattice points
//                (bit order == lexicographic voxel order == torch.unique's row order); self-cleaning.
//   C  objects   radix sort of the object keys, ordered unique + decode, per-segment offsets.
//   D  fuse      fuse_point_clouds_centroid (:88-165).  Every statistical outlier removal (:48-58) the reference
//                would run is one GROUP of points (a stacked pair of clouds, :123-126, or a single cloud, :131 /
//                :147); all groups of the batch are filtered at once (group id = high bits of the grid keys).
//                RT3D_FUSE_BY_INDEX: group = object o of every camera.  RT3D_FUSE_CENTROID: groups follow the
//                classes, then centroids, the greedy per-class matching and the assembly of the fused clouds
//                run on the device as well.
//   E  subtract  subtract_point_clouds_gpu (:168-187) of each frame's workspace (camera clouds
//                concatenated, :248) against its fused objects.
//
// HBM roofline of A (the dominant streaming kernel): (4 + n_obj) * H * W bytes per image in,
// everything else is L2-resident reductions; B reads only occupied bitmap words.
#include <limits.h>

#include <cuda.h>     // CUtensorMap and its enums only: the encoder is fetched through cudaGetDriverEntryPoint
#include <stdlib.h>

#include "knn.cuh"
#include "radix_sort.cuh"
#include "select.cuh"
#include "subtract.cuh"

namespace rt3d {

constexpr int kFbMaxCams = 8;
constexpr int kFbThreads = 256;
constexpr int kFbCellBits = 7;       // batched SOR grid: at most 128 cells per axis
constexpr int kFbCodeBits = 3 * kFbCellBits;

struct FbCam {
    float cx, cy;
    DivBy dfx, dfy;
    Rigid rt;
};

// floor(n / d) for 0 <= n < 2^31 by multiply + shift (Granlund-Montgomery: m = ceil(2^(31+l) / d), l = ceil(log2 d))
struct FastDiv {
    unsigned long long m;
    int sh;
    unsigned d;
    __device__ __forceinline__ unsigned div(unsigned n) const { return (unsigned)(((unsigned long long)n * m) >> sh); }
};
static FastDiv make_fastdiv(unsigned d) {
    FastDiv f;
    int l = 0;
    while ((1ull << l) < d) ++l;
    f.sh = 31 + l;
    f.m = ((1ull << f.sh) + d - 1) / d;
    f.d = d;
    return f;
}

struct FbParams {
    int F, C, H, W, n_obj;
    FastDiv div_c;                  // cameras per frame
    FastDiv div_pj_cpi, div_pj_cx;  // project chunks (kPjChunk tiles of 8 rows x 64 pixels): per image, per tile row
    int cull_cx, cull_cy;           // padded cull table: cells per row / rows per camera
    int word_x, word_c;             // slab-major word id  = ix * word_x + (iyb >> 5) * nzp + iz + word_c  (level 1, run merging)
    int word_cm, word0_cm, k0y_m;   // MODE 2: word_c, word0_c, k0[1] with the rounding constant's bit pattern folded in (fb_pixel_pair)
    int word0_z, word0_x, word0_y, word0_c;  // level-0 word index  = iz * word0_z + ix * word0_x + (iyb >> 5) * word0_y + word0_c  (z-major planes, see bm0_index)
    FastDiv div_nyg;                // slab index -> (x, y group)
    FastDiv div_spi;                // global slab index -> image
    unsigned pj_chunks;             // project chunks of the batch
    FastDiv div_tm_cpi;             // the tensor-map kernel's chunks (kPjChunk stages of kTmRows rows x 64 pixels) per image
    unsigned tm_chunks;
    int force_cold;                 // MODE 2: the cull table reaches beyond the verified fast-division range (plan bit kPlanCullCold)
#ifdef RT3D_PROFILING
    int pj_debug;                   // profiling builds only (RT3D_PJ_DEBUG): 1 = skip the bitmap atomics, 2 = skip the exact path; results are then wrong
#endif
    double sor_cell;                 // outlier-filter grid cell edge: ~3 voxels, see fb_setup
    FbCam cam[kFbMaxCams];
    Box box;
    DivBy div_v;
    float v32;
    int k0[3], nk[3];          // workspace voxel grid inside the crop box
    int nyg, nzp;              // y groups of 32 voxels, z extent padded to a multiple of 32 (bitmap layout)
    int slabs_per_img;         // nk[0] * nyg: one slab = all z words of one (x, y-group)
    int words0_per_img, words1_per_img, tiles_per_img;
    int obits, obias, seg_shift;  // object keys: (seg << seg_shift) | x | y | z, obits per axis, biased
    int sor_k;
    double sor_ratio;
    float sub_t2f;
    double sub_t2;
    int fuse_mode, slots;      // RT3D_FUSE_*; output slots (= outlier-filter groups) per frame
    double fuse_thr;
};

struct FbCounts {     // device-resident element counts of the batch (also returned to the caller)
    int status;       // RT3D_DEV_* bits (must stay first: same place as Ctrl::status)
    int obj_px;       // object pixels (keys appended by A)
    int obj_vox;      // unique object voxels over all segments
    int fused;        // object points surviving the outlier filter
    int ws_vox;       // workspace voxels over all images
    int kept;         // workspace points surviving the subtraction
    int lat_settled;  // outlier-filter queries answered on the bit planes (the rest went through the grid search)
    int ws_filt;      // workspace rows surviving the optional workspace outlier filter (ws_sor)
    int ticket[10];
};

// One outlier-filter group: the points ONE remove_statistical_outlier call of the reference sees -- the stacked
// clouds of a 1-vs-1 class (:123-126) or a single cloud (:131, :147).  Groups of a frame are contiguous in the
// group-ordered point arrays (pd, grp), in the order the reference's loops visit them.
enum { kGrpEmpty = 0, kGrpPair = 1, kGrpCam1 = 2, kGrpCam2 = 3 };
struct PairGeom {
    double lo[3], h, inv_h;
    int dim[3];
    int start, n, n0;  // range in the group-ordered arrays, rows of the first cloud
    double thr;        // outlier threshold (filled by the stats kernel)
    int passthrough;   // RT3D_FUSE_BY_INDEX, camera-0 cloud empty: the reference appends camera 2's cloud unfiltered (:162-163)
    int cols;          // 1: the group has a column table (dim[0] * dim[1] <= kFbColCap)
    int kind, cls;     // kGrp*, class rank (RT3D_FUSE_CENTROID) / object index
    int skip;          // never filtered: a camera-2 cloud of a class camera 1 did not see (:162-163)
    int seg;           // kGrpCam1 / kGrpCam2: the cloud's segment in the voxel arrays (obj_pts rows seg_off[seg]..)
};

constexpr int kFbKnnChunk = 8;    // consecutive (cell-sorted) queries per ticket: neighbours share their candidate list
constexpr int kFbColCap = 16384;  // (x, y) columns of a pair's SOR grid that get a start-index table


// --------------------------------------------------------------------------------- A: project
// `a / b` in the arithmetic flavour MODE (see DivBy): 0 IEEE, 1 reciprocal multiply, 2 verified FMA-corrected IEEE
template <int MODE>
__device__ __forceinline__ float div_m(float a, const DivBy &d) {
    if (MODE == 2) {
        const float q = __fmul_rn(a, d.inv_b);
        return __fmaf_rn(__fmaf_rn(-q, d.b, a), d.inv_b, q);
    }
    return MODE == 1 ? __fmul_rn(a, d.inv_b) : __fdiv_rn(a, d.b);
}
// round-half-even voxel index: cvt.rni.s32.f32 rounds exactly like torch.round / rintf
template <int MODE>
__device__ __forceinline__ int voxel_m(float p, const DivBy &div_v) { return __float2int_rn(div_m<MODE>(p, div_v)); }

// Depths the verified fast division may see.  fb_setup only selects MODE 2 when |cx|, |cy| are zero or in
// [2^-20, 2^20] and the image is at most 8192 wide/high, so a non-zero |u - cx| lies in [2^-43, 2^21] and
// every dividend (u - cx) * d stays inside the exhaustively verified exponent range [-100, 100).
__device__ __forceinline__ bool depth_in_fast_range(float d) { return d >= 0x1p-50f && d <= 0x1p70f; }

constexpr unsigned kNoBit = 0xffffffffu;

// Level-0 occupancy words are stored plane by plane (z major).  Inside a plane the x index runs fastest: the patch of
// a mostly horizontal surface that a warp commits at once spans several x at one or two (y group, z), so its words
// share 32-byte sectors -- one L2 atomic request per sector instead of one per word (RT3D_BM_XFAST 0: y group fastest).
#ifndef RT3D_BM_XFAST
#define RT3D_BM_XFAST 1
#endif
// How a thread's words reach the two bitmap levels:
//   RT3D_PJ_L1_RED 1 (default)  fire-and-forget reductions into both levels (level 1 once per run of equal level-1 words of a
//                               thread).  No return values, so nothing is parked across patches and the patch loop can stay
//                               rolled: a quarter of the code (the unrolled kernel lost a third of its patch time to
//                               instruction fetch: ncu stall_no_inst 2.1 warps per issue).  Measured 0.856 ms at 256 frames.
//   RT3D_PJ_L1_RED 0            level 0 by atomics whose return values tell who found the word empty and owes the level-1
//                               bit (half the reductions, but the return values are parked in registers and every parked
//                               atomic holds a scoreboard; needs RT3D_PJ_UNROLL 1): 0.879 ms.
#ifndef RT3D_PJ_L1_RED
#define RT3D_PJ_L1_RED 1
#endif
#ifndef RT3D_PJ_UNROLL
#define RT3D_PJ_UNROLL (RT3D_PJ_L1_RED ? 0 : 1)
#endif
// (ix, iz biased by the crop box's first voxel when word0_c is added, unbiased otherwise)
__device__ __forceinline__ unsigned bm0_index(const FbParams &P, unsigned ix, unsigned iyg, unsigned iz) {
#if RT3D_BM_XFAST
    return iz * (unsigned)P.word0_z + iyg * (unsigned)P.word0_y + ix;
#else
    return iz * (unsigned)P.word0_z + ix * (unsigned)P.word0_x + iyg;
#endif
}

// Depth-interval cull (built once per configuration by fb_cull_build_kernel): for a 4x4 pixel cell of
// camera c, the set of depths whose back-projected, transformed point can fall inside the crop box is an
// interval (p = d * R ray + t is linear in d).  The table holds a conservative fp32 enclosure of it over the
// cell's 16 rays, widened by far more than the fp32 rounding error of the exact path, as raw float bits
// {lo, hi - lo}: for positive finite floats the bit patterns order like the values, and zero, negatives,
// +-inf and NaN all fall outside [lo, hi] as unsigned integers -- so (bits(d) - lo) <= span is the validity
// test (:26) and the box pre-test in two integer instructions.  Survivors still take the EXACT arithmetic
// and the exact closed-interval box test; the table only removes pixels that cannot pass it.
struct __align__(8) CullCell { unsigned lo, span; };
constexpr int kCullCell = 4;   // cell edge in pixels

// The table is padded to whole project tiles (8 rows x 64 pixels); padding cells reject everything, so the
// project kernel needs no image-edge tests on the workspace path.
__global__ void __launch_bounds__(256) fb_cull_build_kernel(const __grid_constant__ FbParams P, CullCell *cull, unsigned *meta) {
    const int cells_x = P.cull_cx, cells_y = P.cull_cy;
    const long long per_cam = (long long)cells_x * cells_y, total = per_cam * P.C;
    const double e = 0x1p-17;   // relative slack: ~100x the worst-case fp32 error of the exact path
    for (long long idx = blockIdx.x * 256ll + threadIdx.x; idx < total; idx += gridDim.x * 256ll) {
        const int cam = (int)(idx / per_cam);
        const long long c = idx - cam * per_cam;
        const int cyi = (int)(c / cells_x), cxi = (int)(c - (long long)cyi * cells_x);
        const FbCam &K = P.cam[cam];
        double lo = DBL_MAX, hi = -DBL_MAX;
        for (int dv = 0; dv < kCullCell; ++dv) {
            const int v = cyi * kCullCell + dv;
            if (v >= P.H) break;
            for (int du = 0; du < kCullCell; ++du) {
                const int u = cxi * kCullCell + du;
                if (u >= P.W) break;
                const double ray[3] = {((double)u - (double)K.cx) / (double)K.dfx.b, ((double)v - (double)K.cy) / (double)K.dfy.b, 1.0};
                double a = 0.0, b = DBL_MAX;   // this pixel's admissible depths
                bool empty = false;
                for (int k = 0; k < 3; ++k) {
                    const double r0 = (double)K.rt.R[3 * k] * ray[0], r1 = (double)K.rt.R[3 * k + 1] * ray[1], r2 = (double)K.rt.R[3 * k + 2];
                    const double rk = r0 + r1 + r2, sk = fabs(r0) + fabs(r1) + fabs(r2);
                    const double tk = (double)K.rt.t[k], blo = (double)P.box.lo[k], bhi = (double)P.box.hi[k];
                    const double ck = fabs(tk) + fmax(fabs(blo), fabs(bhi)) + 1e-30;
                    // d * (rk - e sk) <= bhi + e ck - tk   and   -d * (rk + e sk) <= -(blo - e ck - tk)
                    const double ca[2] = {rk - e * sk, -(rk + e * sk)};
                    const double cb[2] = {bhi + e * ck - tk, -(blo - e * ck - tk)};
                    for (int s = 0; s < 2; ++s) {
                        if (ca[s] > 0.0) b = fmin(b, cb[s] / ca[s]);
                        else if (ca[s] < 0.0) a = fmax(a, cb[s] / ca[s]);
                        else if (cb[s] < 0.0) empty = true;
                    }
                }
                if (!empty && a <= b) { lo = fmin(lo, a); hi = fmax(hi, b); }
            }
        }
        CullCell cell;
        cell.lo = 0xffffffffu; cell.span = 0u;   // nothing can pass (a NaN with that exact pattern fails the box test)
        if (lo <= hi) {
            float lf = __double2float_rd(lo * (1.0 - 1e-6)), hf = __double2float_ru(fmin(hi * (1.0 + 1e-6), 3.0e38));
            unsigned lb = lf > 0.0f ? __float_as_uint(lf) : 1u;          // smallest positive float
            unsigned hb = hf > 0.0f ? __float_as_uint(hf) : 1u;
            if (hb > 0x7f7fffffu) hb = 0x7f7fffffu;
            if (lb < 1u) lb = 1u;
            if (hb >= lb) { cell.lo = lb; cell.span = hb - lb; atomicMax(&meta[0], hb); }
        }
        cull[idx] = cell;
    }
}

// cold path (a candidate whose depth lies outside the fast-division range): plain IEEE divisions, out of line
__device__ __noinline__ uint4 fb_cold_pixel(const FbParams *Pp, const FbCam *Kp, float fu, float fvc, float d) {
    const FbParams &P = *Pp;
    const FbCam &K = *Kp;
    float x = __fdiv_rn(__fmul_rn(__fsub_rn(fu, K.cx), d), K.dfx.b), y = __fdiv_rn(__fmul_rn(fvc, d), K.dfy.b), z = d;
    K.rt.apply(x, y, z, x, y, z);
    uint4 r = make_uint4(kNoBit, 0u, kNoBit, 0u);   // {slab-major word id, bits, level-0 word index, -}
    if (depth_valid(d) && P.box.contains(x, y, z)) {
        const unsigned ix = (unsigned)(__float2int_rn(__fdiv_rn(x, P.div_v.b)) - P.k0[0]), iy = (unsigned)(__float2int_rn(__fdiv_rn(y, P.div_v.b)) - P.k0[1]),
                       iz = (unsigned)(__float2int_rn(__fdiv_rn(z, P.div_v.b)) - P.k0[2]);
        r.x = (ix * (unsigned)P.nyg + (iy >> 5)) * (unsigned)P.nzp + iz;
        r.y = 1u << (iy & 31);
        r.z = bm0_index(P, ix, iy >> 5, iz);
    }
    return r;
}

// one object pixel (mask set, depth valid): packed (segment | voxel) key.  Object clouds are not cropped
// (:300-302), so their coordinates are unbounded and MODE 2 falls back to plain IEEE division here.
template <int MODE>
__device__ __noinline__ void fb_obj_emit(const FbParams *Pp, const FbCam *Kp, unsigned u, unsigned v, float d, unsigned long long seg,
                                         unsigned long long *okeys, int pos, int ocap, FbCounts *cnt) {
    const FbParams &P = *Pp;
    const FbCam &K = *Kp;
    const float ax = __fmul_rn(__fsub_rn((float)u, K.cx), d), ay = __fmul_rn(__fsub_rn((float)v, K.cy), d);
    float x, y, z = d;
    if (MODE == 1) { x = __fmul_rn(ax, K.dfx.inv_b); y = __fmul_rn(ay, K.dfy.inv_b); }
    else { x = __fdiv_rn(ax, K.dfx.b); y = __fdiv_rn(ay, K.dfy.b); }
    K.rt.apply(x, y, z, x, y, z);
    int kx, ky, kz;
    if (MODE == 1) { kx = __float2int_rn(__fmul_rn(x, P.div_v.inv_b)); ky = __float2int_rn(__fmul_rn(y, P.div_v.inv_b)); kz = __float2int_rn(__fmul_rn(z, P.div_v.inv_b)); }
    else { kx = __float2int_rn(__fdiv_rn(x, P.div_v.b)); ky = __float2int_rn(__fdiv_rn(y, P.div_v.b)); kz = __float2int_rn(__fdiv_rn(z, P.div_v.b)); }
    const int lim = 2 * P.obias;
    const int bx = kx + P.obias, by = ky + P.obias, bz = kz + P.obias;
    if (bx < 0 || by < 0 || bz < 0 || bx >= lim || by >= lim || bz >= lim)
        atomicOr(&cnt->status, RT3D_DEV_KEY_RANGE);  // beyond max_range: reported, key clamped
    const unsigned long long key = (seg << P.seg_shift) | ((unsigned long long)min(max(bx, 0), lim - 1) << (2 * P.obits)) |
                                   ((unsigned long long)min(max(by, 0), lim - 1) << P.obits) | (unsigned long long)min(max(bz, 0), lim - 1);
    if (pos < ocap) okeys[pos] = key; else atomicOr(&cnt->status, RT3D_DEV_CAPACITY);
}

// Commit a thread's (word, bits) entries of one patch to the two-level occupancy bitmap.  Executed by the
// whole warp.  The level-1 bit of a word is set by whoever finds the level-0 word still empty (the atomic's
// return value), so level 1 costs one reduction per distinct word of the image, not per point.
//   DEDUP 0  every entry goes straight to memory
//   DEDUP 1  lanes holding the same (word, bits) pair elect one of them (match_any on the pair)
//   DEDUP 2  lanes holding the same word OR their bits together by pointer jumping over the match_any
//            group (log steps, no per-group loop), the lowest lane commits
constexpr unsigned kFastLoBits = 0x26800000u;   // bits(2^-50): a candidate below it is positive (the cull says so) but tiny
constexpr unsigned kFastHiBits = 0x62800000u;   // bits(2^70)
constexpr float kRoundMagic = 12582912.0f;      // 1.5 * 2^23
constexpr unsigned kRoundMagicBits = 0x4B400000u;
// Two neighbouring workspace pixels at once with Blackwell's packed fp32 instructions (FFMA2 / FMUL2 / FADD2:
// each does two IEEE round-to-nearest operations per issue slot, numerically identical per component to the
// scalar intrinsics), i.e. the same results as two fb_pixel calls at half the floating-point instruction count.
__device__ __forceinline__ float2 f2(float a) { return make_float2(a, a); }
template <int MODE>
__device__ __forceinline__ float2 div_m2(float2 a, const DivBy &d) {
    if (MODE == 2) {   // q = a * inv;  q + (a - q b) * inv, the verified FMA-corrected IEEE quotient
        const float2 q = __fmul2_rn(a, f2(d.inv_b));
        return __ffma2_rn(__ffma2_rn(q, f2(-d.b), a), f2(d.inv_b), q);
    }
    if (MODE == 1) return __fmul2_rn(a, f2(d.inv_b));
    return make_float2(__fdiv_rn(a.x, d.b), __fdiv_rn(a.y, d.b));
}
// closed-interval box test without short-circuit evaluation (`&&` compiles to a divergent branch per pixel, with its
// reconvergence barrier and reloaded bounds; six compares chained on one predicate are cheaper than the branch)
__device__ __forceinline__ bool box_contains_nobranch(const Box &b, float x, float y, float z) {
    return (x >= b.lo[0]) & (x <= b.hi[0]) & (y >= b.lo[1]) & (y <= b.hi[1]) & (z >= b.lo[2]) & (z <= b.hi[2]);
}
template <int MODE>
__device__ __forceinline__ void fb_pixel_pair(const FbParams &P, const FbCam &K, const CullCell cell, float fu, float fvc, float da, float db,
                                              unsigned &wa, unsigned &w0a, unsigned &ba, unsigned &wb, unsigned &w0b, unsigned &bb, bool &cold) {
    const bool canda = (__float_as_uint(da) - cell.lo) <= cell.span, candb = (__float_as_uint(db) - cell.lo) <= cell.span;
    const float2 d = make_float2(da, db);
    float2 x = div_m2<MODE>(__fmul2_rn(__fadd2_rn(make_float2(fu, fu + 1.0f), f2(-K.cx)), d), K.dfx);
    float2 y = div_m2<MODE>(__fmul2_rn(f2(fvc), d), K.dfy);
    // torch.mm's rounding: fma chain over k, then a rounded + t (Rigid::apply, two points at a time)
    const float *R = K.rt.R, *t = K.rt.t;
    const float2 px = __fadd2_rn(__ffma2_rn(d, f2(R[2]), __ffma2_rn(y, f2(R[1]), __fmul2_rn(x, f2(R[0])))), f2(t[0]));
    const float2 py = __fadd2_rn(__ffma2_rn(d, f2(R[5]), __ffma2_rn(y, f2(R[4]), __fmul2_rn(x, f2(R[3])))), f2(t[1]));
    const float2 pz = __fadd2_rn(__ffma2_rn(d, f2(R[8]), __ffma2_rn(y, f2(R[7]), __fmul2_rn(x, f2(R[6])))), f2(t[2]));
    const bool inba = canda & box_contains_nobranch(P.box, px.x, py.x, pz.x), inbb = candb & box_contains_nobranch(P.box, px.y, py.y, pz.y);
    float2 qx = div_m2<MODE>(px, P.div_v), qy = div_m2<MODE>(py, P.div_v), qz = div_m2<MODE>(pz, P.div_v);
    // Round-half-even voxel indices.  MODE 2 (only selected when every in-box index is below 2^21 in magnitude) rounds by
    // adding 1.5 * 2^23 -- the sum's ulp is 1, so the addition rounds the quotient exactly like cvt.rni / torch.round --
    // and reads the integer off the mantissa: packed adds on the FMA pipe instead of six conversions on the XU pipe, with
    // the constant's bit pattern folded into the index constants (word0_cm, k0y_m).  Outside the box the indices are
    // garbage either way and masked by the box test.
    int ixa, iya, iza, ixb, iyb, izb;
    unsigned c0;
    if (MODE == 2) {
        qx = __fadd2_rn(qx, f2(kRoundMagic)); qy = __fadd2_rn(qy, f2(kRoundMagic)); qz = __fadd2_rn(qz, f2(kRoundMagic));
        ixa = __float_as_int(qx.x); iya = __float_as_int(qy.x) - P.k0y_m; iza = __float_as_int(qz.x);
        ixb = __float_as_int(qx.y); iyb = __float_as_int(qy.y) - P.k0y_m; izb = __float_as_int(qz.y);
        c0 = (unsigned)P.word0_cm;
    } else {
        ixa = __float2int_rn(qx.x); iya = __float2int_rn(qy.x) - P.k0[1]; iza = __float2int_rn(qz.x);
        ixb = __float2int_rn(qx.y); iyb = __float2int_rn(qy.y) - P.k0[1]; izb = __float2int_rn(qz.y);
        c0 = (unsigned)P.word0_c;
    }
    const unsigned c1 = MODE == 2 ? (unsigned)P.word_cm : (unsigned)P.word_c;   // slab-major level-1 ids
    wa = (unsigned)ixa * (unsigned)P.word_x + (unsigned)(iya >> 5) * (unsigned)P.nzp + (unsigned)iza + c1;
    wb = (unsigned)ixb * (unsigned)P.word_x + (unsigned)(iyb >> 5) * (unsigned)P.nzp + (unsigned)izb + c1;
    const unsigned idxa = bm0_index(P, (unsigned)ixa, (unsigned)(iya >> 5), (unsigned)iza) + c0;
    const unsigned idxb = bm0_index(P, (unsigned)ixb, (unsigned)(iyb >> 5), (unsigned)izb) + c0;
    w0a = inba ? idxa : kNoBit; ba = 1u << (iya & 31);
    w0b = inbb ? idxb : kNoBit; bb = 1u << (iyb & 31);
    if (MODE == 2) cold = cold | (canda & (__float_as_uint(da) < kFastLoBits)) | (candb & (__float_as_uint(db) < kFastLoBits));   // rare: below the verified range
}

// A thread's (word, bits) entries go straight to the level-0 bitmap with atomicOr; whoever finds the word
// still empty (the atomic's return value) owes the level-1 bit, so level 1 costs one reduction per distinct
// word of an image, not per point.  The return values are NOT awaited here: they are parked in `Pending`
// and examined after the next patch's atomics have been issued (or at the end of the kernel), so the L2
// round trip overlaps the next patch's arithmetic.
struct Pending {
    unsigned w0, w1, w2, w3, o0, o1, o2, o3;   // w = level-1 id of the word; o = old level-0 word, or non-zero when nothing is owed
    unsigned *b1;
};
// predicated atomics as single instructions (no branch around them)
__device__ __forceinline__ unsigned atom_or_if(bool p, unsigned *addr, unsigned v, unsigned otherwise) {
    unsigned old = otherwise;
    asm volatile("{ .reg .pred q; setp.ne.u32 q, %3, 0; @q atom.global.or.b32 %0, [%1], %2; }" : "+r"(old) : "l"(addr), "r"(v), "r"((unsigned)p) : "memory");
    return old;
}
__device__ __forceinline__ void red_or_if(bool p, unsigned *addr, unsigned v) {
    asm volatile("{ .reg .pred q; setp.ne.u32 q, %2, 0; @q red.global.or.b32 [%0], %1; }" ::"l"(addr), "r"(v), "r"((unsigned)p) : "memory");
}
__device__ __forceinline__ void fb_flush_pending(const Pending &p) {
    // a level-0 word is found empty once per image and word: nearly always nothing is owed
    if (__all_sync(__activemask(), (p.o0 != 0u) & (p.o1 != 0u) & (p.o2 != 0u) & (p.o3 != 0u))) return;
    red_or_if(p.o0 == 0u, p.b1 + (p.w0 >> 5), 1u << (p.w0 & 31));
    red_or_if(p.o1 == 0u, p.b1 + (p.w1 >> 5), 1u << (p.w1 & 31));
    red_or_if(p.o2 == 0u, p.b1 + (p.w2 >> 5), 1u << (p.w2 & 31));
    red_or_if(p.o3 == 0u, p.b1 + (p.w3 >> 5), 1u << (p.w3 & 31));
}

// ---- asynchronous global -> shared copies (LDGSTS): each lane copies, and later reads, only its own slots,
// so completion is tracked per thread with commit / wait groups and the warp never synchronises
__device__ __forceinline__ unsigned smem_addr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
// 16-byte streaming copy; `valid` false copies nothing and zero-fills the slot (a zero-size copy reads nothing, so
// the address may lie outside the image)
__device__ __forceinline__ void cp_async_16_stream_zfill(bool valid, unsigned dst, const void *src, unsigned long long policy) {
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2, %3;" ::"r"(dst), "l"(src), "r"(valid ? 16u : 0u), "l"(policy) : "memory");
}
__device__ __forceinline__ void cp_async_8(unsigned dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_4(unsigned dst, const void *src, unsigned src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
// ---- mbarrier + TMA (cp.async.bulk.tensor) primitives of the tensor-map pipeline (fb_project_tma_kernel)
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Bounded wait (the house rule: a wedged pipeline becomes a status bit, never a hang).  try_wait suspends the thread
// for a hardware-defined time slice per attempt, so the bound is seconds, far beyond any legitimate wait.
constexpr unsigned long long kMbarWaitLimitNs = 2000000000ull;   // 2 s
__device__ __forceinline__ bool mbar_try_wait(unsigned bar, unsigned parity) {
    unsigned done;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    return done != 0u;
}
__device__ __forceinline__ bool mbar_test_wait(unsigned bar, unsigned parity) {   // non-blocking probe
    unsigned done;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    return done != 0u;
}
__device__ __forceinline__ bool mbar_wait(unsigned bar, unsigned parity) {
    if (mbar_try_wait(bar, parity)) return true;
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
#pragma unroll 1
        for (int i = 0; i < 64; ++i)
            if (mbar_try_wait(bar, parity)) return true;
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > kMbarWaitLimitNs) return false;
    }
}
__device__ __forceinline__ void tma_load_4d(unsigned dst, const void *tmap, int c0, int c1, int c2, int c3, unsigned bar, unsigned long long policy) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4, %5}], [%6], %7;"
                 ::"r"(dst), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar), "l"(policy) : "memory");
}
__device__ __forceinline__ void tma_load_3d(unsigned dst, const void *tmap, int c0, int c1, int c2, unsigned bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Warp tile = 8 rows x 64 pixels of one image, as four 8-row x 16-pixel patches; lane (r, cg) owns the four
// consecutive pixels (one 128-bit shared-memory read) at row r, column group cg of each patch.  Compact 2-D
// patches keep a warp's occupancy atomics inside a handful of bitmap sectors.  A staged tile holds, per patch q,
//   depth   [q][lane] float4                      (2048 B)
//   masks   [o][q][lane] 4 bytes, o < kPjStageMasks (512 B per staged instance mask)
//   cull    the 4 x 4-pixel cells under the tile   (layout differs between the two kernels, see CQ below)
// Two kernels fill such stages and share the per-tile arithmetic (pj_process_tile):
//   fb_project_tma_kernel   image width a multiple of 16: a producer warp moves whole tiles with three TMA tensor
//                           copies (cp.async.bulk.tensor) into a block-level ring, consumer warps take tiles from it
//   fb_project_kernel       any width (a multiple of 4): every warp stages its own next tile with per-lane cp.async
constexpr int kPjRows = 8, kPjCols = 64, kPjChunk = 4, kPjStageMasks = 4;
constexpr int kPjDepthBytes = 2048, kPjMaskBytes = 512;
constexpr unsigned kPjCamBytes = sizeof(FbCam) * kFbMaxCams;

struct PjShared {       // per-kernel constants of the tile arithmetic
    const FbCam *cam;   // camera table in shared memory
    unsigned *bm0, *bm1;
    unsigned long long *okeys;
    int ocap;
    FbCounts *cnt;
    const unsigned char *masks;   // global masks (instance masks beyond the staged ones are read from here)
    bool force_cold;
};

// One tile (omask: bit o clear = staged instance mask o has no pixel in the tile): sd / sc / smk = this lane's view of the staged depth, cull cells and mask words (patch q at sd[32 RS q],
// sc[CQ q], smk[128 RS o + 32 RS q]; RS = 8-row tiles per staged block: a patch of a 16-row block holds 16 rows);
// (img, v, ub) = image, this lane's row and its first column (patch 0).
template <int MODE, int CQ, int RS = 1>
__device__ __forceinline__ void pj_process_tile(const FbParams &P, const PjShared &S, const float4 *sd, const CullCell *sc, const unsigned *smk,
                                                unsigned img_, unsigned v_, unsigned ub_, Pending &pend, unsigned omask = ~0u) {
    constexpr int DQ = 32 * RS, MQ = 32 * RS, MO = (kPjMaskBytes / 4) * RS;   // strides: depth patch (float4), mask patch, mask plane (words)
    const int lane = lane_id();
    const unsigned HW = (unsigned)P.H * (unsigned)P.W;
    const int nm = P.n_obj < kPjStageMasks ? P.n_obj : kPjStageMasks;
    unsigned anym = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float4 d = sd[q * DQ];
        const CullCell c = sc[q * CQ];
        const bool any = (__float_as_uint(d.x) - c.lo) <= c.span || (__float_as_uint(d.y) - c.lo) <= c.span ||
                         (__float_as_uint(d.z) - c.lo) <= c.span || (__float_as_uint(d.w) - c.lo) <= c.span;
        anym |= any ? (1u << q) : 0u;
    }
    // ---- workspace: exact math for the candidates, one atomic per distinct bitmap word of a thread
#ifdef RT3D_PROFILING
    if (P.pj_debug == 2) anym = 0;
#endif
    if (__ballot_sync(kFullMask, anym != 0) != 0) {   // warp-uniform
        unsigned img = img_, v = v_, ub = ub_;
        asm volatile("" : "+r"(img), "+r"(v), "+r"(ub));   // keep everything below inside this branch (no speculative hoisting)
        const unsigned cam = img - P.div_c.div(img) * P.div_c.d;
        const FbCam K = S.cam[cam];
        const float fvc = __fsub_rn((float)v, K.cy);
        unsigned *b0 = S.bm0 + (size_t)img * P.words0_per_img, *b1 = S.bm1 + (size_t)img * P.words1_per_img;
        const bool force_cold = MODE == 2 && S.force_cold;   // warp-uniform (a plan bit, see rt3d_frame_batch_init)
        // RT3D_PJ_UNROLL 1: one copy of the patch code per patch.  In a rolled loop the scoreboard of the parked atomics
        // is drained at the loop head, i.e. each patch waits for its predecessor's L2 round trips before its own
        // arithmetic starts; builds without parked return values (RT3D_PJ_L1_RED) can roll the loop (a quarter of the code).
#if RT3D_PJ_UNROLL
#pragma unroll
#else
#pragma unroll 1
#endif
        for (int q = 0; q < 4; ++q) {
            if (__ballot_sync(kFullMask, (anym >> q) & 1u) == 0) continue;   // warp-uniform
            if ((anym >> q) & 1u) {
                unsigned w0, w1, w2, w3, t0, t1, t2, t3, a0, a1, a2, a3;   // level-1 ids, bits, level-0 word indices (kNoBit: none)
                const float4 d = sd[q * DQ];      // the stage is re-read: cheaper than holding 16 depths in registers
                const CullCell c = sc[q * CQ];    // (a lane outside the image never gets here: its `any` bit is clear)
                const float fu0 = (float)(ub + 16u * q);
                bool cold = force_cold;
                fb_pixel_pair<MODE>(P, K, c, fu0, fvc, d.x, d.y, w0, a0, t0, w1, a1, t1, cold);
                fb_pixel_pair<MODE>(P, K, c, fu0 + 2.0f, fvc, d.z, d.w, w2, a2, t2, w3, a3, t3, cold);
                if (MODE == 2 && cold) {   // redo the patch with plain IEEE division, out of line
                    const FbCam *Ks = &S.cam[cam];
                    uint4 e;
                    e = fb_cold_pixel(&P, Ks, fu0, fvc, d.x); w0 = e.x; t0 = e.y; a0 = e.z;
                    e = fb_cold_pixel(&P, Ks, fu0 + 1.0f, fvc, d.y); w1 = e.x; t1 = e.y; a1 = e.z;
                    e = fb_cold_pixel(&P, Ks, fu0 + 2.0f, fvc, d.z); w2 = e.x; t2 = e.y; a2 = e.z;
                    e = fb_cold_pixel(&P, Ks, fu0 + 3.0f, fvc, d.w); w3 = e.x; t3 = e.y; a3 = e.z;
                }
                // runs of equal words inside the thread: the run's last slot carries the OR of its bits
#ifdef RT3D_PROFILING
                if (P.pj_debug == 1) { a0 = a1 = a2 = a3 = kNoBit; }
#endif
                const bool k0 = a0 != kNoBit && a0 != a1;
                t1 |= (a1 == a0) ? t0 : 0u;
                const bool k1 = a1 != kNoBit && a1 != a2;
                t2 |= (a2 == a1) ? t1 : 0u;
                const bool k2 = a2 != kNoBit && a2 != a3;
                t3 |= (a3 == a2) ? t2 : 0u;
                const bool k3 = a3 != kNoBit;
#if RT3D_PJ_L1_RED
                // Both levels by fire-and-forget reductions: nothing comes back, so no warp ever waits for an L2 round
                // trip.  Level 1 (one bit per level-0 word) is then set by every committed word, not only by the first
                // toucher; runs of equal level-1 words inside the thread are merged like the level-0 runs.
                red_or_if(k0, b0 + a0, t0);
                red_or_if(k1, b0 + a1, t1);
                red_or_if(k2, b0 + a2, t2);
                red_or_if(k3, b0 + a3, t3);
                unsigned m0 = k0 ? 1u << (w0 & 31) : 0u, m1 = k1 ? 1u << (w1 & 31) : 0u, m2 = k2 ? 1u << (w2 & 31) : 0u, m3 = k3 ? 1u << (w3 & 31) : 0u;
                const unsigned l0 = w0 >> 5, l1 = w1 >> 5, l2 = w2 >> 5, l3 = w3 >> 5;
                const bool j0 = l0 != l1;
                m1 |= j0 ? 0u : m0;
                const bool j1 = l1 != l2;
                m2 |= j1 ? 0u : m1;
                const bool j2 = l2 != l3;
                m3 |= j2 ? 0u : m2;
                red_or_if(j0 && m0 != 0u, b1 + l0, m0);
                red_or_if(j1 && m1 != 0u, b1 + l1, m1);
                red_or_if(j2 && m2 != 0u, b1 + l2, m2);
                red_or_if(m3 != 0u, b1 + l3, m3);
#else
                // the previous patch's return values have had this patch's arithmetic to arrive: settle them,
                // then park the new atomics' return values in the same registers (no copy, no wait)
                fb_flush_pending(pend);
                pend.w0 = w0; pend.w1 = w1; pend.w2 = w2; pend.w3 = w3; pend.b1 = b1;
                pend.o0 = atom_or_if(k0, b0 + a0, t0, 1u);
                pend.o1 = atom_or_if(k1, b0 + a1, t1, 1u);
                pend.o2 = atom_or_if(k2, b0 + a2, t2, 1u);
                pend.o3 = atom_or_if(k3, b0 + a3, t3, 1u);
#endif
            }
        }
    }

    // ---- instance masks: append packed (segment | voxel) keys; their order is irrelevant (sorted next)
    // (two loops: the staged masks come from shared memory only.  In one loop the scoreboard of the global loads of
    // the un-staged case joins the wait of every mask word -- and it is the scoreboard of the parked bitmap atomics.)
    auto mask_keys = [&](int o, const unsigned (&mw)[4]) {
        if (__ballot_sync(kFullMask, (mw[0] | mw[1] | mw[2] | mw[3]) != 0u) == 0) return;
        unsigned img = img_, v = v_, ub = ub_;
        asm volatile("" : "+r"(img), "+r"(v), "+r"(ub));
        unsigned m16 = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (mw[q] == 0u) continue;
            const float4 d = sd[q * DQ];
            const float dq[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (((mw[q] >> (8 * j)) & 0xffu) && depth_valid(dq[j])) m16 |= 1u << (q * 4 + j);
        }
        if (__ballot_sync(kFullMask, m16 != 0) == 0) return;
        const int c = __popc(m16);
        const int incl = warp_incl_scan(c);
        const int tot = __shfl_sync(kFullMask, incl, 31);
        int pos0 = 0;
        if (lane == 31) pos0 = atomicAdd(&S.cnt->obj_px, tot);
        int pos = __shfl_sync(kFullMask, pos0, 31) + incl - c;
        if (m16) {
            const unsigned frame = P.div_c.div(img), cam = img - frame * P.div_c.d;
            const FbCam *Ks = &S.cam[cam];
            const unsigned long long seg = (unsigned long long)((frame * P.n_obj + o) * P.C + cam);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (!((m16 >> (4 * q)) & 15u)) continue;
                const float4 d = sd[q * DQ];
                const float dq[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (m16 & (1u << (q * 4 + j))) fb_obj_emit<MODE>(&P, Ks, ub + 16u * q + j, v, dq[j], seg, S.okeys, pos++, S.ocap, S.cnt);
            }
        }
    };
    for (int o = 0; o < nm; ++o) {
        if (!((omask >> o) & 1u)) continue;   // the caller found the staged plane all zero (warp-uniform)
        unsigned mw[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) mw[q] = smk[o * MO + q * MQ];
        mask_keys(o, mw);
    }
    for (int o = nm; o < P.n_obj; ++o) {   // more instance masks than the stage holds: straight from global memory
        unsigned mw[4];
        const bool row_ok = v_ < (unsigned)P.H;
        const unsigned char *mrow = S.masks + ((size_t)img_ * P.n_obj + o) * HW + (size_t)(row_ok ? v_ : 0u) * (unsigned)P.W;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const unsigned u0 = ub_ + 16u * q;
            mw[q] = (row_ok && u0 < (unsigned)P.W) ? ldg_stream_u32(mrow + u0) : 0u;
        }
        mask_keys(o, mw);
    }
}

// ----------------------------------------------------------------- A.1 tensor-map (TMA) producer / consumer pipeline
// Block = kTmConsumers consumer warps + 1 producer warp.  Every consumer warp owns a private ring of kTmDepth stages
// (a stage = kTmRS vertically adjacent tiles); lane w of the producer warp is the dedicated producer of consumer warp w.
// The producer lanes run in SIMT lock-step: each polls (non-blocking test_wait) the empty barrier of its ring's next
// stage and, when it is free, takes the next stage of its current chunk (kPjChunk horizontally adjacent stages per
// ticket), posts its descriptor and the TMA tensor copies -- depth box {16 px, kTmRows rows, 4 patches}, the cull cells
// under it, the staged instance masks -- that land on the stage's full barrier.  Pixels outside the image are zero-filled
// by the TMA unit (depth 0 = invalid, mask 0): no edge code anywhere.  A consumer that is busy with a dense stage delays
// nobody: its producer lane simply finds no free stage and the tickets it does not take go to the other warps.  A stage
// is only ever waited for by the warp that consumed its previous phase, so the barriers' phase parity cannot alias
// (with one shared ring a warp can be handed slot j + S while slot j's bytes are still in flight, and then sees the
// unfinished phase as "completed").
#ifndef RT3D_TM_CONSUMERS
#define RT3D_TM_CONSUMERS 8
#endif
#ifndef RT3D_TM_DEPTH
#define RT3D_TM_DEPTH 3   // stages per ring with 8-row stages; 16-row stages take 2 (the same bytes in flight)
#endif
#ifndef RT3D_TM_BLOCKS
#define RT3D_TM_BLOCKS 2
#endif
// A stage holds kTmRS vertically adjacent tiles (one tensor copy each for depth, cull cells and every staged mask, whatever
// the box height): with 16-row stages the producer posts, and a consumer waits for, half as many stages per pixel -- the
// producer warp was busy 80 % of the time with 8-row stages and the consumers waited for it (ncu: 21 % of all stall
// samples at the full barrier while the kernel ran at 0.37 of the HBM roofline).
#ifndef RT3D_TM_ROWS
#define RT3D_TM_ROWS 16
#endif
constexpr int kTmRS = RT3D_TM_ROWS / kPjRows, kTmRows = RT3D_TM_ROWS;
static_assert(kTmRS == 1 || kTmRS == 2, "stages of 8 or 16 rows");
constexpr int kTmConsumers = RT3D_TM_CONSUMERS, kTmDepth = kTmRS == 2 && RT3D_TM_DEPTH > 2 ? 2 : RT3D_TM_DEPTH, kTmBlocksPerSM = RT3D_TM_BLOCKS;
constexpr int kTmStages = kTmConsumers * kTmDepth;
constexpr int kTmThreads = 32 * (kTmConsumers + 1);
constexpr int kTmDepthBytes = kPjDepthBytes * kTmRS, kTmMaskBytes = kPjMaskBytes * kTmRS;
constexpr int kTmCullBytes = 256 * kTmRS;   // 2 rows of 16 cells per 8 image rows
static_assert(kTmConsumers <= 32, "one producer lane per consumer warp");
__host__ __device__ inline unsigned tm_stage_bytes(int n_obj) {
    const unsigned nm = (unsigned)(n_obj < kPjStageMasks ? n_obj : kPjStageMasks);
    return kTmDepthBytes + kTmCullBytes + kTmMaskBytes * nm;     // every part a multiple of 128 B (TMA destination alignment)
}
// dynamic shared memory: [camera table | pad to 128][stages: warp-major][descriptors: kTmStages x uint4][barriers: full, empty: 2 x kTmStages x 8 B]
__host__ __device__ inline unsigned tm_ring_off() { return (kPjCamBytes + 127u) & ~127u; }
__host__ __device__ inline unsigned tm_smem_bytes(int n_obj) { return tm_ring_off() + kTmStages * (tm_stage_bytes(n_obj) + 16u + 16u); }

struct TmMaps {   // the three tensor maps of a launch (host-encoded, passed by value: they must live in param / const space)
    CUtensorMap depth, masks, cull;
};

template <int MODE>
__global__ void __launch_bounds__(kTmThreads, kTmBlocksPerSM) fb_project_tma_kernel(const __grid_constant__ FbParams P, const __grid_constant__ TmMaps M,
                                                                                    const unsigned char *__restrict__ masks,
                                                                                    unsigned *bm0, unsigned *bm1,
                                                                                    unsigned long long *okeys, int ocap, FbCounts *cnt) {
    extern __shared__ __align__(128) unsigned char pj_smem[];
    FbCam *s_cam = reinterpret_cast<FbCam *>(pj_smem);
    const unsigned stage_bytes = tm_stage_bytes(P.n_obj);
    unsigned char *ring = pj_smem + tm_ring_off();
    uint4 *s_desc = reinterpret_cast<uint4 *>(ring + kTmStages * stage_bytes);                 // {img, v0, x0, -}; img == kNoBit: no more tiles
    unsigned long long *s_bar = reinterpret_cast<unsigned long long *>(s_desc + kTmStages);   // full[kTmStages], empty[kTmStages]
    const unsigned ring_s = smem_addr(ring), full_s = smem_addr(s_bar), empty_s = full_s + 8u * kTmStages;
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    if (threadIdx.x < P.C) s_cam[threadIdx.x] = P.cam[threadIdx.x];
    if (threadIdx.x == 0) {
        for (int st = 0; st < kTmStages; ++st) { mbar_init(full_s + 8u * st, 1u); mbar_init(empty_s + 8u * st, 1u); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int nm = P.n_obj < kPjStageMasks ? P.n_obj : kPjStageMasks;

    if (warp == kTmConsumers) {
        // ---------------- producers: lane w feeds consumer warp w
        if (lane >= kTmConsumers) return;
        asm volatile("prefetch.tensormap [%0];" ::"l"(&M.depth) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&M.cull) : "memory");
        if (nm > 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&M.masks) : "memory");
        unsigned long long pol_stream;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_stream));
        const unsigned first = (unsigned)lane * kTmDepth;      // this ring's first stage
        unsigned k = 0;                                         // tiles posted to this ring
        // Chunk tickets come from one global counter (dynamic: a chunk over an object costs several times a chunk of
        // background, and with a few frames per batch a static deal leaves whole blocks idle at the end -- measured at
        // 8 HD1080 frames x 12 masks: 0.67 vs 0.62 ms; at 256 HD720 frames the two are equal).  Every lane requests its
        // next ticket when it starts a chunk, i.e. four tiles before it needs it.
        unsigned next = atomicAdd(&cnt->ticket[4], 1);
        unsigned left = 0, img = 0, cam = 0, v0 = 0, x0 = 0;
        unsigned long long t_idle = 0;                          // start of the current run of failed polls (0: none)
        unsigned polls = 0;
        for (;;) {
            const unsigned stage = first + k % kTmDepth;
            if (!mbar_test_wait(empty_s + 8u * stage, ((k / kTmDepth) & 1u) ^ 1u)) {
                // the consumer has not freed the stage yet: poll again (bounded: a wedged consumer becomes a status bit)
                if ((++polls & 1023u) == 0u) {
                    unsigned long long now;
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                    if (t_idle == 0) t_idle = now;
                    else if (now - t_idle > kMbarWaitLimitNs) { atomicOr(&cnt->status, RT3D_DEV_LOOKBACK); return; }
                }
                continue;
            }
            t_idle = 0;
            if (left == 0) {
                if (next >= P.tm_chunks) break;
                const unsigned chunk = next;
                next = atomicAdd(&cnt->ticket[4], 1);
                img = P.div_tm_cpi.div(chunk);
                const unsigned c_in = chunk - img * P.div_tm_cpi.d;
                const unsigned ty = P.div_pj_cx.div(c_in), cx = c_in - ty * P.div_pj_cx.d;
                cam = img - P.div_c.div(img) * P.div_c.d;
                v0 = ty * kTmRows;
                x0 = cx * (kPjChunk * kPjCols);
                left = kPjChunk;
            }
            const unsigned x = x0;
            x0 += kPjCols; --left;
            if (x >= (unsigned)P.W) { left = 0; continue; }        // the chunk's tail lies outside the image
            s_desc[stage] = make_uint4(img, v0, x, 0u);
            const unsigned dst = ring_s + stage * stage_bytes, bar = full_s + 8u * stage;
            mbar_expect_tx(bar, stage_bytes);      // (release: the descriptor is visible to whoever sees the phase complete)
            tma_load_4d(dst, &M.depth, 0, (int)v0, (int)(x >> 4), (int)img, bar, pol_stream);
            tma_load_3d(dst + kTmDepthBytes, &M.cull, (int)(x >> 1), (int)(v0 >> 2), (int)cam, bar);
            for (int o = 0; o < nm; ++o)
                tma_load_4d(dst + kTmDepthBytes + kTmCullBytes + o * kTmMaskBytes, &M.masks, 0, (int)v0, (int)(x >> 4), (int)(img * P.n_obj + o), bar,
                            pol_stream);
            ++k;
        }
        {   // end marker (the free stage was found just above)
            const unsigned stage = first + k % kTmDepth;
            s_desc[stage] = make_uint4(kNoBit, 0u, 0u, 0u);
            mbar_arrive(full_s + 8u * stage);
        }
        return;
    }

    // ---------------- consumers
    PjShared S;
    S.cam = s_cam; S.bm0 = bm0; S.bm1 = bm1; S.okeys = okeys; S.ocap = ocap; S.cnt = cnt; S.masks = masks;
    // MODE 2: a cull interval reaching beyond the verified fast-division range sends every candidate down the cold path
    // (decided by rt3d_frame_batch_init and carried in the parameters: as a word loaded here, the comparison is re-done inside
    // the tile loop on the scoreboard of the parked bitmap atomics -- every patch then waited for its predecessor's atomics
    // half-way through its arithmetic: 20 % of all stall samples on that one ISETP)
    S.force_cold = MODE == 2 && P.force_cold != 0;
    const unsigned r = (unsigned)lane >> 2, cg = (unsigned)lane & 3u;
    Pending pend;
    pend.o0 = pend.o1 = pend.o2 = pend.o3 = 1u; pend.w0 = pend.w1 = pend.w2 = pend.w3 = 0u; pend.b1 = bm1;
    const unsigned first = (unsigned)warp * kTmDepth;
    for (unsigned k = 0;; ++k) {
        const unsigned stage = first + k % kTmDepth;
        if (!mbar_wait(full_s + 8u * stage, (k / kTmDepth) & 1u)) { atomicOr(&cnt->status, RT3D_DEV_LOOKBACK); break; }
        const uint4 t = s_desc[stage];
        if (t.x == kNoBit) break;
        const unsigned char *sb = ring + stage * stage_bytes;
        // which staged instance masks have a pixel anywhere in the stage: two 128-bit reads per lane and plane, whatever
        // the pixel layout (most stages hold none: the per-tile mask code is then skipped altogether)
        unsigned omask = 0u;
        for (int o = 0; o < nm; ++o) {
            const uint4 *mp = reinterpret_cast<const uint4 *>(sb + kTmDepthBytes + kTmCullBytes + o * kTmMaskBytes) + lane;
            uint4 a = mp[0];
            unsigned any = a.x | a.y | a.z | a.w;
            if (kTmRS == 2) { a = mp[32]; any |= a.x | a.y | a.z | a.w; }
            if (__any_sync(kFullMask, any != 0u)) omask |= 1u << o;
        }
#pragma unroll 1
        for (unsigned hh = 0; hh < (unsigned)kTmRS; ++hh) {   // the stage's 8-row tiles, top to bottom (rows below the image: nothing staged)
            if (t.y + hh * kPjRows >= (unsigned)P.H) break;
            pj_process_tile<MODE, 4, kTmRS>(P, S, reinterpret_cast<const float4 *>(sb) + hh * 32u + lane,
                                            reinterpret_cast<const CullCell *>(sb + kTmDepthBytes) + (2u * hh + (r >> 2)) * 16u + cg,
                                            reinterpret_cast<const unsigned *>(sb + kTmDepthBytes + kTmCullBytes) + hh * 32u + lane, t.x,
                                            t.y + hh * kPjRows + r, t.z + cg * 4u, pend, omask);
        }
        __syncwarp();                                   // every lane has finished reading the stage
        if (lane == 0) mbar_arrive(empty_s + 8u * stage);
    }
    fb_flush_pending(pend);
}

// ----------------------------------------------------------------- A.2 per-warp cp.async pipeline (any width)
// Every warp runs its own pipeline: an atomic ticket hands out chunks of kPjChunk horizontally adjacent tiles (one
// index decode per chunk), and tile n+1's depth / cull cells / mask words are in flight (cp.async into the warp's
// other smem stage) while tile n is processed -- no block-wide barrier, no per-tile exposure to DRAM latency.
constexpr int kPjStages = 2, kPjBlocksPerSM = 2, kPjThreads = 256;
constexpr int kPjCullBytes = 1024;   // one private copy of its cell per lane and patch
__host__ __device__ inline unsigned pj_stage_bytes(int n_obj) {
    const unsigned nm = (unsigned)(n_obj < kPjStageMasks ? n_obj : kPjStageMasks);
    return kPjDepthBytes + kPjCullBytes + kPjMaskBytes * nm;
}
__host__ __device__ inline unsigned pj_warp_bytes(int n_obj) { return kPjStages * pj_stage_bytes(n_obj); }

struct PjTile {   // one tile, lane view; img == kNoBit: no tile
    unsigned img, v, ub;   // image, this lane's row and first column (patch 0)
    unsigned dpix, cpos;   // this lane's first pixel in the depth batch, its first cell in the cull table
};

template <int MODE>
__global__ void __launch_bounds__(kPjThreads, kPjBlocksPerSM) fb_project_kernel(const __grid_constant__ FbParams P, const float *__restrict__ depth,
                                                                                const unsigned char *__restrict__ masks,
                                                                                const CullCell *__restrict__ cull,
                                                                                unsigned *bm0, unsigned *bm1, unsigned long long *okeys, int ocap,
                                                                                FbCounts *cnt) {
    extern __shared__ __align__(128) unsigned char pj_smem[];
    FbCam *s_cam = reinterpret_cast<FbCam *>(pj_smem);
    if (threadIdx.x < P.C) s_cam[threadIdx.x] = P.cam[threadIdx.x];
    __syncthreads();
    const unsigned HW = (unsigned)P.H * (unsigned)P.W;
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    const unsigned r = (unsigned)lane >> 2, cg = (unsigned)lane & 3u;
    const int nm = P.n_obj < kPjStageMasks ? P.n_obj : kPjStageMasks;
    const unsigned stage_bytes = pj_stage_bytes(P.n_obj);
    unsigned char *wsm = pj_smem + kPjCamBytes + (size_t)warp * pj_warp_bytes(P.n_obj);   // the warp's stages
    const unsigned wsm_s = smem_addr(wsm);
    unsigned long long pol_stream;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_stream));
    PjShared S;
    S.cam = s_cam; S.bm0 = bm0; S.bm1 = bm1; S.okeys = okeys; S.ocap = ocap; S.cnt = cnt; S.masks = masks;
    // (decided by rt3d_frame_batch_init and carried in the parameters: as a word loaded here, the comparison is re-done inside
    // the tile loop on the scoreboard of the parked bitmap atomics -- every patch then waited for its predecessor's atomics
    // half-way through its arithmetic: 20 % of all stall samples on that one ISETP)
    S.force_cold = MODE == 2 && P.force_cold != 0;

    // Lanes outside the image (right edge of a padded chunk, rows below H) get depth 0 = invalid and mask 0: a cull
    // cell may straddle the last image row (H need not be a multiple of 4), so a stale slot could pass it.
    auto issue = [&](const PjTile &t, unsigned stage) {
        const unsigned sb = wsm_s + stage * stage_bytes;
        const bool row_ok = t.v < (unsigned)P.H;
        const float *dsrc = depth + t.dpix;
        const CullCell *csrc = cull + t.cpos;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            cp_async_16_stream_zfill(row_ok && t.ub + 16u * q < (unsigned)P.W, sb + (unsigned)lane * 16u + q * 512u, dsrc + 16 * q, pol_stream);
            cp_async_8(sb + kPjDepthBytes + (unsigned)lane * 8u + q * 256u, csrc + 4 * q);
        }
        if (nm > 0) {
            const unsigned char *msrc = masks + (size_t)t.img * P.n_obj * HW + (t.v * (unsigned)P.W + t.ub);
            for (int o = 0; o < nm; ++o, msrc += HW) {
#pragma unroll
                for (int q = 0; q < 4; ++q)   // (a zero-size copy reads nothing: the address may lie outside the image)
                    cp_async_4(sb + kPjDepthBytes + kPjCullBytes + o * kPjMaskBytes + (unsigned)lane * 4u + q * 128u, msrc + 16 * q,
                               (row_ok && t.ub + 16u * q < (unsigned)P.W) ? 4u : 0u);
            }
        }
    };

    // chunked tickets; the next chunk's ticket is requested one chunk ahead
    unsigned tk = 0, left = 0;
    if (lane == 0) tk = atomicAdd(&cnt->ticket[4], 1);
    PjTile feed;   // issue side
    feed.img = kNoBit; feed.v = 0; feed.ub = 0; feed.dpix = 0; feed.cpos = 0;
    auto fetch = [&]() {   // advances `feed` to the next tile (img = kNoBit once the tickets are exhausted)
        if (left == 0) {
            const unsigned ticket = __shfl_sync(kFullMask, tk, 0);
            if (ticket >= P.pj_chunks) { feed.img = kNoBit; return; }
            if (lane == 0) tk = atomicAdd(&cnt->ticket[4], 1);
            const unsigned chunk = ticket;   // (a scattered visiting order was measured: no gain, 64 frames: 0.340 vs 0.332 ms)
            feed.img = P.div_pj_cpi.div(chunk);
            const unsigned c_in = chunk - feed.img * P.div_pj_cpi.d;
            const unsigned ty = P.div_pj_cx.div(c_in), cx = c_in - ty * P.div_pj_cx.d;
            const unsigned cam = feed.img - P.div_c.div(feed.img) * P.div_c.d;
            feed.v = ty * kPjRows + r;
            feed.ub = cx * (kPjChunk * kPjCols) + cg * 4u;
            feed.dpix = feed.img * HW + feed.v * (unsigned)P.W + feed.ub;    // < 2^32 (checked on the host)
            feed.cpos = (cam * (unsigned)P.cull_cy + (feed.v >> 2)) * (unsigned)P.cull_cx + (feed.ub >> 2);
            left = kPjChunk - 1;
        } else {
            feed.ub += kPjCols; feed.dpix += kPjCols; feed.cpos += kPjCols / kCullCell;
            --left;
        }
    };
    PjTile ring[kPjStages];   // ring[0] = the tile being processed, the others are in flight
#pragma unroll
    for (int st = 0; st < kPjStages - 1; ++st) {
        fetch();
        ring[st] = feed;
        if (feed.img != kNoBit) issue(feed, (unsigned)st);
        cp_async_commit();
    }
    Pending pend;
    pend.o0 = pend.o1 = pend.o2 = pend.o3 = 1u; pend.w0 = pend.w1 = pend.w2 = pend.w3 = 0u; pend.b1 = bm1;
    unsigned stage = 0, fill = kPjStages - 1;   // stage of ring[0]; stage the next copy goes to
    while (ring[0].img != kNoBit) {
        if (feed.img != kNoBit) fetch();
        ring[kPjStages - 1] = feed;
        if (feed.img != kNoBit) issue(feed, fill);
        cp_async_commit();
        cp_async_wait<kPjStages - 1>();
        const unsigned char *sb = wsm + stage * stage_bytes;
        pj_process_tile<MODE, 32>(P, S, reinterpret_cast<const float4 *>(sb) + lane, reinterpret_cast<const CullCell *>(sb + kPjDepthBytes) + lane,
                                  reinterpret_cast<const unsigned *>(sb + kPjDepthBytes + kPjCullBytes) + lane, ring[0].img, ring[0].v, ring[0].ub, pend);
#pragma unroll
        for (int st = 0; st < kPjStages - 1; ++st) ring[st] = ring[st + 1];
        stage = stage + 1 == kPjStages ? 0u : stage + 1;
        fill = fill + 1 == kPjStages ? 0u : fill + 1;
    }
    cp_async_wait<0>();
    fb_flush_pending(pend);
}

// an overflowing object stream is reported (RT3D_DEV_CAPACITY) and truncated, never read past its buffer
__global__ void fb_clamp_kernel(FbCounts *cnt, int ocap) {
    if (cnt->obj_px > ocap) cnt->obj_px = ocap;
}

// --------------------------------------------------------------------------------- B: extract
// The occupancy bitmaps become each image's sorted unique lattice points.  A slab = all z words of one
// (x, y group of 32); inside a slab lane y owns voxel row y, so emitting slabs in (x, y group) order and rows
// in (y, z) order is exactly the lexicographic (x, y, z) order of torch.unique.  Three embarrassingly
// parallel steps (no inter-block waiting anywhere):
//   count    one warp per slab: voxels of the slab (level-1 words name the occupied level-0 words)
//   offsets  one block per image: exclusive scan of its slab counts; one more block: scan of the image totals
//   emit     one warp per slab: write its rows at the slab's offset, clearing both bitmap levels on the way
constexpr int kExtCap = 256;   // occupied words of a slab held in shared memory (more -> direct global walk)

__global__ void __launch_bounds__(kFbThreads) fb_ws_count_kernel(const __grid_constant__ FbParams P, const unsigned *__restrict__ bm0,
                                                                 const unsigned *__restrict__ bm1, int *__restrict__ slab_cnt) {
    const int lane = lane_id();
    const int nl1 = P.nzp >> 5;   // level-1 words per slab
    const long long nslab = (long long)P.F * P.C * P.slabs_per_img;
    const long long w0 = (blockIdx.x * (long long)kFbThreads + threadIdx.x) >> 5, nw = ((long long)gridDim.x * kFbThreads) >> 5;
    if (nl1 <= 16) {   // (the usual case: a crop box up to 512 voxels tall) two slabs per warp, 16 lanes each: no idle half
        const int half = lane >> 4, hl = lane & 15;
        const long long npair = (nslab + 1) >> 1;
        for (long long pg = w0; pg < npair; pg += nw) {
            const long long sg = 2 * pg + half;
            int c = 0;
            if (sg < nslab) {
                const int img = (int)P.div_spi.div((unsigned)sg), slab = (int)((unsigned)sg - (unsigned)img * P.div_spi.d);
                const unsigned sx = P.div_nyg.div((unsigned)slab), sy = (unsigned)slab - sx * P.div_nyg.d;
                const unsigned *b0 = bm0 + (size_t)img * P.words0_per_img + bm0_index(P, sx, sy, 0u);
                unsigned w1 = hl < nl1 ? bm1[(size_t)img * P.words1_per_img + (size_t)slab * nl1 + hl] : 0u;
                for (; w1; w1 &= w1 - 1) c += __popc(b0[(size_t)(hl * 32 + (__ffs(w1) - 1)) * P.slabs_per_img]);
            }
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) c += __shfl_xor_sync(kFullMask, c, o);
            if (hl == 0 && sg < nslab) slab_cnt[sg] = c;
        }
        return;
    }
    for (long long sg = w0; sg < nslab; sg += nw) {
        const int img = (int)P.div_spi.div((unsigned)sg), slab = (int)((unsigned)sg - (unsigned)img * P.div_spi.d);   // (a 64-bit division costs ~100 instructions)
        const unsigned sx = P.div_nyg.div((unsigned)slab), sy = (unsigned)slab - sx * P.div_nyg.d;
        const unsigned *b0 = bm0 + (size_t)img * P.words0_per_img + bm0_index(P, sx, sy, 0u);   // word of plane z: b0[z * slabs_per_img]
        const unsigned *b1 = bm1 + (size_t)img * P.words1_per_img + (size_t)slab * nl1;
        int c = 0;
        for (int l0 = 0; l0 < nl1; l0 += 32) {
            // lane l reads level-1 word l; every lane then walks ITS word's occupied level-0 words
            const int l = l0 + lane;
            unsigned w1 = l < nl1 ? b1[l] : 0u;
            for (; w1; w1 &= w1 - 1) c += __popc(b0[(size_t)(l * 32 + (__ffs(w1) - 1)) * P.slabs_per_img]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(kFullMask, c, o);
        if (lane == 0) slab_cnt[sg] = c;
    }
}

// one block per image: slab_cnt -> offsets relative to the image's first row (in place), image total
__global__ void __launch_bounds__(kFbThreads) fb_ws_slab_scan_kernel(const __grid_constant__ FbParams P, int *slab_cnt, int *img_total) {
    __shared__ int s_tmp[8];
    __shared__ int s_run;
    const int img = blockIdx.x;
    int *cntp = slab_cnt + (size_t)img * P.slabs_per_img;
    if (threadIdx.x == 0) s_run = 0;
    __syncthreads();
    for (int base = 0; base < P.slabs_per_img; base += kFbThreads) {
        const int i = base + threadIdx.x;
        const int v = i < P.slabs_per_img ? cntp[i] : 0;
        int total;
        const int excl = block_excl_scan_256(v, s_tmp, total);
        const int run = s_run;
        if (i < P.slabs_per_img) cntp[i] = run + excl;
        __syncthreads();
        if (threadIdx.x == 0) s_run = run + total;
        __syncthreads();
    }
    if (threadIdx.x == 0) img_total[img] = s_run;
}

// one block: ws_off = exclusive scan of the image totals
__global__ void __launch_bounds__(kFbThreads) fb_ws_img_scan_kernel(const int *__restrict__ img_total, int nimg, int *ws_off, FbCounts *cnt) {
    __shared__ int s_tmp[8];
    __shared__ int s_run;
    if (threadIdx.x == 0) s_run = 0;
    __syncthreads();
    for (int base = 0; base < nimg; base += kFbThreads) {
        const int i = base + threadIdx.x;
        const int v = i < nimg ? img_total[i] : 0;
        int total;
        const int excl = block_excl_scan_256(v, s_tmp, total);
        const int run = s_run;
        if (i < nimg) ws_off[i] = run + excl;
        __syncthreads();
        if (threadIdx.x == 0) s_run = run + total;
        __syncthreads();
    }
    if (threadIdx.x == 0) { ws_off[nimg] = s_run; cnt->ws_vox = s_run; }
}

__global__ void __launch_bounds__(kFbThreads) fb_ws_emit_kernel(const __grid_constant__ FbParams P, const unsigned *__restrict__ bm0, const unsigned *__restrict__ bm1,
                                                                const int *__restrict__ slab_off, const int *__restrict__ ws_off,
                                                                float *ws_xyz, int ws_cap, FbCounts *cnt) {
    __shared__ unsigned short s_z[kFbThreads / 32][kExtCap];
    __shared__ unsigned s_w[kFbThreads / 32][kExtCap];
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    const int nl1 = P.nzp >> 5;
    const long long nslab = (long long)P.F * P.C * P.slabs_per_img;
    const long long w0 = (blockIdx.x * (long long)kFbThreads + threadIdx.x) >> 5, nw = ((long long)gridDim.x * kFbThreads) >> 5;
    for (long long sg = w0; sg < nslab; sg += nw) {
        const int img = (int)P.div_spi.div((unsigned)sg), slab = (int)((unsigned)sg - (unsigned)img * P.div_spi.d);   // (a 64-bit division costs ~100 instructions)
        const unsigned ixs = P.div_nyg.div((unsigned)slab), iyg = (unsigned)slab - ixs * P.div_nyg.d;
        const unsigned *b0 = bm0 + (size_t)img * P.words0_per_img + bm0_index(P, ixs, iyg, 0u);   // word of plane z: b0[z * slabs_per_img]
        const unsigned *b1 = bm1 + (size_t)img * P.words1_per_img + (size_t)slab * nl1;
        const size_t zs = (size_t)P.slabs_per_img;
        // occupied words of the slab, ascending z
        int nnz = 0;
        for (int l0 = 0; l0 < nl1; l0 += 32) {
            const int l = l0 + lane;
            unsigned w1 = l < nl1 ? b1[l] : 0u;
            const int c = __popc(w1);
            const int incl = warp_incl_scan(c);
            int at = nnz + incl - c;
            for (; w1; w1 &= w1 - 1, ++at)
                if (at < kExtCap) s_z[warp][at] = (unsigned short)(l * 32 + (__ffs(w1) - 1));
            nnz += __shfl_sync(kFullMask, incl, 31);
        }
        if (nnz == 0) continue;   // warp-uniform
        __syncwarp();
        const bool staged = nnz <= kExtCap;
        int mine = 0;   // voxels of row y = lane
        if (staged) {
            for (int i = lane; i < nnz; i += 32) { const int z = s_z[warp][i]; s_w[warp][i] = b0[z * zs]; }
            __syncwarp();
            for (int i = 0; i < nnz; ++i) mine += (s_w[warp][i] >> lane) & 1u;
        } else {
            for (int z = 0; z < P.nzp; ++z) mine += (b0[z * zs] >> lane) & 1u;   // every lane reads the same word
        }
        const int incl = warp_incl_scan(mine);
        long long pos = (long long)ws_off[img] + slab_off[sg] + incl - mine;
        const float fx = __fmul_rn((float)((int)ixs + P.k0[0]), P.v32);
        const float fy = __fmul_rn((float)((int)(iyg * 32u + (unsigned)lane) + P.k0[1]), P.v32);
        if (staged) {
            for (int i = 0; i < nnz && mine; ++i) {
                if (!((s_w[warp][i] >> lane) & 1u)) continue;
                if (pos < ws_cap) {
                    float *o = ws_xyz + 3 * pos;
                    o[0] = fx; o[1] = fy; o[2] = __fmul_rn((float)((int)s_z[warp][i] + P.k0[2]), P.v32);
                } else {
                    atomicOr(&cnt->status, RT3D_DEV_CAPACITY);
                }
                ++pos;
            }
        } else {
            for (int z = 0; z < P.nzp; ++z) {
                if (!((b0[z * zs] >> lane) & 1u)) continue;
                if (pos < ws_cap) {
                    float *o = ws_xyz + 3 * pos;
                    o[0] = fx; o[1] = fy; o[2] = __fmul_rn((float)(z + P.k0[2]), P.v32);
                } else {
                    atomicOr(&cnt->status, RT3D_DEV_CAPACITY);
                }
                ++pos;
            }
        }
        __syncwarp();
    }
}

// One block per image, after the rows are out: both bitmap levels go back to all-zero for the next batch.
// Scattered 4-byte clears from the emit kernel cost more than the extraction itself (measured: 0.34 of 0.59 ms
// at 64 frame pairs); here level 1 is OR-reduced into a mask of occupied z planes while it is zeroed with
// contiguous stores, and only those planes of the z-major level 0 are zeroed, again contiguously.
constexpr int kClearThreads = 512;
__global__ void __launch_bounds__(kClearThreads) fb_bm_clear_kernel(const __grid_constant__ FbParams P, unsigned *bm0, unsigned *bm1) {
    __shared__ unsigned s_planes[2048];   // nzp <= 65535 -> at most 2048 level-1 words per slab
    const int nl1 = P.nzp >> 5;
    const int img = blockIdx.x;
    unsigned *b1 = bm1 + (size_t)img * P.words1_per_img;
    unsigned *b0 = bm0 + (size_t)img * P.words0_per_img;
    for (int l = threadIdx.x; l < nl1; l += kClearThreads) s_planes[l] = 0u;
    __syncthreads();
    const int n1 = P.slabs_per_img * nl1;
    {   // level 1: word i belongs to z group i % nl1 (kept incrementally: a modulo per word costs more than the word)
        const int step = kClearThreads % nl1;
        int rem = threadIdx.x % nl1;
        for (int i = threadIdx.x; i < n1; i += kClearThreads) {
            const unsigned w = b1[i];
            if (w) { atomicOr(&s_planes[rem], w); b1[i] = 0u; }
            rem += step;
            if (rem >= nl1) rem -= nl1;
        }
    }
    __syncthreads();
    // level 0: runs of consecutive occupied planes are contiguous memory -- zeroed with 16-byte stores between a scalar
    // head and tail (a plane of slabs_per_img words is no multiple of 16 bytes)
    const size_t S = (size_t)P.slabs_per_img;
    int z = 0;
    const int nz = P.nzp;
    while (z < nz) {   // block-uniform scan of the plane mask
        if (!((s_planes[z >> 5] >> (z & 31)) & 1u)) { ++z; continue; }
        int z1 = z + 1;
        while (z1 < nz && ((s_planes[z1 >> 5] >> (z1 & 31)) & 1u)) ++z1;
        unsigned *p0 = b0 + (size_t)z * S, *p1 = b0 + (size_t)z1 * S;
        unsigned *a0 = reinterpret_cast<unsigned *>(((uintptr_t)p0 + 15u) & ~(uintptr_t)15u);
        unsigned *a1 = reinterpret_cast<unsigned *>((uintptr_t)p1 & ~(uintptr_t)15u);
        if (a0 >= a1) {
            for (unsigned *q = p0 + threadIdx.x; q < p1; q += kClearThreads) *q = 0u;
        } else {
            if (threadIdx.x < (unsigned)(a0 - p0)) p0[threadIdx.x] = 0u;
            if (threadIdx.x < (unsigned)(p1 - a1)) a1[threadIdx.x] = 0u;
            uint4 *v0 = reinterpret_cast<uint4 *>(a0), *v1 = reinterpret_cast<uint4 *>(a1);
            for (uint4 *q = v0 + threadIdx.x; q < v1; q += kClearThreads) *q = make_uint4(0u, 0u, 0u, 0u);
        }
        z = z1;
    }
}

// --------------------------------------------------------------------------------- C: objects
struct FbObjUniqueOp {
    using Payload = unsigned long long;
    const unsigned long long *keys;
    FbParams P;
    float *obj_pts;
    unsigned long long *obj_key;
    __device__ __forceinline__ unsigned eval4(long long i0, long long n, unsigned long long (&p)[4]) const {
        unsigned keep = 0;
        unsigned long long prev = i0 > 0 ? keys[i0 - 1] : 0ull;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const long long i = i0 + j;
            if (i >= n) break;
            const unsigned long long cur = keys[i];
            if (i == 0 || cur != prev) { p[j] = cur; keep |= 1u << j; }
            prev = cur;
        }
        return keep;
    }
    __device__ __forceinline__ void store(long long pos, const unsigned long long &key) const {
        const unsigned long long fm = (1ull << P.obits) - 1ull;
        const int kz = (int)(key & fm) - P.obias, ky = (int)((key >> P.obits) & fm) - P.obias,
                  kx = (int)((key >> (2 * P.obits)) & fm) - P.obias;
        float *o = obj_pts + 3 * pos;
        o[0] = __fmul_rn((float)kx, P.v32); o[1] = __fmul_rn((float)ky, P.v32); o[2] = __fmul_rn((float)kz, P.v32);
        obj_key[pos] = key;
    }
};

// off[s] = first index whose (key >> shift) >= s, for s = 0..nseg (off[nseg] = n)
__global__ void fb_offsets_u64_kernel(const unsigned long long *__restrict__ keys, const int *n_dev, int shift, int nseg, int *off) {
    const int n = *n_dev;
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s <= nseg; s += gridDim.x * blockDim.x) {
        int lo = 0, hi = n;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if ((keys[mid] >> shift) < (unsigned long long)s) lo = mid + 1; else hi = mid; }
        off[s] = lo;
    }
}
// same over an ascending int array scaled by `div` (segment of element i = vals[i] / div)
__global__ void fb_offsets_i32_kernel(const int *__restrict__ vals, const int *n_dev, int div, int nseg, int *off) {
    const int n = *n_dev;
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s <= nseg; s += gridDim.x * blockDim.x) {
        int lo = 0, hi = n;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (vals[mid] / div < s) lo = mid + 1; else hi = mid; }
        off[s] = lo;
    }
}

// --------------------------------------------------------------------------------- D: fuse
// ---- D.0 grouping.  One thread per frame turns the frame's segments (camera clouds, sizes known only now) into
// outlier-filter groups, in the order the reference's loops would create them:
//   RT3D_FUSE_BY_INDEX   group o = the clouds of object o of every camera, stacked in camera order (:123-126)
//   RT3D_FUSE_CENTROID   classes by ascending rank; a class with exactly one non-empty cloud per camera -> one
//                        kGrpPair group (:121-126); any other class -> one group per camera-1 cloud (:131), then one
//                        per camera-2 cloud (:147: a candidate's filtered copy is needed whenever it is matched)
// A frame's groups occupy the frame's own range of the voxel arrays, so the regrouping is a permutation inside it.
__global__ void __launch_bounds__(128) fb_fuse_plan_kernel(const FbParams P, const int *__restrict__ seg_off, const int *__restrict__ class_rank,
                                                           PairGeom *geom, int *seg_dst, int *seg_group, FbCounts *cnt) {
    const int f = blockIdx.x * 128 + threadIdx.x;
    if (f >= P.F) return;
    const int nseg_f = P.n_obj * P.C, s0 = f * nseg_f, g0 = f * P.slots;
    int cursor = seg_off[s0], g = g0;
    PairGeom e;
    e.lo[0] = e.lo[1] = e.lo[2] = 0.0; e.h = 1.0; e.inv_h = 1.0; e.dim[0] = e.dim[1] = e.dim[2] = 1;
    e.thr = 0.0; e.passthrough = 0; e.cols = 0; e.skip = 0; e.seg = -1;
    if (P.fuse_mode == RT3D_FUSE_BY_INDEX) {
        for (int o = 0; o < P.n_obj; ++o, ++g) {
            const int sa = s0 + o * P.C;
            e.start = seg_off[sa]; e.n = seg_off[sa + P.C] - e.start; e.n0 = seg_off[sa + 1] - e.start;
            e.kind = e.n > 0 ? kGrpPair : kGrpEmpty; e.cls = o; e.passthrough = (e.n0 == 0) ? 1 : 0;
            for (int c = 0; c < P.C; ++c) { seg_dst[sa + c] = seg_off[sa + c]; seg_group[sa + c] = g; }
            geom[g] = e;
        }
        return;
    }
    // segment of (object o, camera c) = s0 + o * 2 + c; class_rank is [F][C][n_obj].  A rank outside [0, slots) is
    // reported and clamped, so that every point still belongs to a group.
    auto rank_of = [&](int c, int o) {
        const int rk = class_rank[f * nseg_f + c * P.n_obj + o];
        if (rk < 0 || rk >= P.slots) atomicOr(&cnt->status, RT3D_DEV_KEY_RANGE);
        return min(max(rk, 0), P.slots - 1);
    };
    for (int r = 0; r < P.slots; ++r) {
        int n1 = 0, n2 = 0, a1 = -1, a2 = -1;
        for (int o = 0; o < P.n_obj; ++o) {
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int sg = s0 + o * 2 + c;
                if (rank_of(c, o) != r || seg_off[sg + 1] == seg_off[sg]) continue;
                if (c == 0) { ++n1; a1 = sg; } else { ++n2; a2 = sg; }
            }
        }
        if (n1 == 0 && n2 == 0) continue;
        if (n1 == 1 && n2 == 1) {
            const int na = seg_off[a1 + 1] - seg_off[a1], nb = seg_off[a2 + 1] - seg_off[a2];
            e.start = cursor; e.n = na + nb; e.n0 = na; e.kind = kGrpPair; e.cls = r; e.skip = 0; e.seg = -1;
            seg_dst[a1] = cursor; seg_dst[a2] = cursor + na; seg_group[a1] = g; seg_group[a2] = g;
            geom[g++] = e;
            cursor += na + nb;
            continue;
        }
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            for (int o = 0; o < P.n_obj; ++o) {
                const int sg = s0 + o * 2 + c;
                if (rank_of(c, o) != r || seg_off[sg + 1] == seg_off[sg]) continue;
                const int n = seg_off[sg + 1] - seg_off[sg];
                e.start = cursor; e.n = n; e.n0 = n; e.kind = c == 0 ? kGrpCam1 : kGrpCam2; e.cls = r;
                e.skip = (c == 1 && n1 == 0) ? 1 : 0; e.seg = sg;
                seg_dst[sg] = cursor; seg_group[sg] = g;
                geom[g++] = e;
                cursor += n;
            }
        }
    }
    e.start = cursor; e.n = 0; e.n0 = 0; e.kind = kGrpEmpty; e.cls = -1; e.skip = 1; e.seg = -1;
    for (; g < g0 + P.slots; ++g) geom[g] = e;
}

// ---- D.1 regroup: voxel arrays (segment order) -> fp64 working copy in group order + the group of every point
__global__ void __launch_bounds__(kFbThreads) fb_fuse_gather_kernel(const FbParams P, const float *__restrict__ obj_pts,
                                                                    const unsigned long long *__restrict__ obj_key, const int *n_dev,
                                                                    const int *__restrict__ seg_off, const int *__restrict__ seg_dst,
                                                                    const int *__restrict__ seg_group, double *__restrict__ pd,
                                                                    int *__restrict__ grp, unsigned long long *__restrict__ gcode) {
    const int n = *n_dev;
    const unsigned long long code_mask = (1ull << P.seg_shift) - 1ull;
    const long long stride = (long long)gridDim.x * kFbThreads;
    for (long long i = (long long)blockIdx.x * kFbThreads + threadIdx.x; i < n; i += stride) {
        const int sg = (int)(obj_key[i] >> P.seg_shift);
        const long long j = (long long)seg_dst[sg] + (i - seg_off[sg]);
        pd[3 * j] = (double)obj_pts[3 * i]; pd[3 * j + 1] = (double)obj_pts[3 * i + 1]; pd[3 * j + 2] = (double)obj_pts[3 * i + 2];
        grp[j] = seg_group[sg];
        gcode[j] = obj_key[i] & code_mask;      // biased voxel indices x | y | z, obits each
    }
}

// one block per group: bounding box, grid geometry
__global__ void __launch_bounds__(128) fb_sor_prepare_kernel(const FbParams P, const double *__restrict__ pd, PairGeom *geom) {
    __shared__ double s_lo[3][4], s_hi[3][4];
    const int p = blockIdx.x;
    const int a = geom[p].start, b = a + geom[p].n;
    if (a == b || geom[p].skip) return;   // block-uniform
    double lo[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, hi[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
    for (int i = a + threadIdx.x; i < b; i += 128) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const double v = pd[3 * (long long)i + k];
            lo[k] = fmin(lo[k], v); hi[k] = fmax(hi[k], v);
        }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[k] = fmin(lo[k], __shfl_xor_sync(kFullMask, lo[k], o));
            hi[k] = fmax(hi[k], __shfl_xor_sync(kFullMask, hi[k], o));
        }
        if (lane_id() == 0) { s_lo[k][threadIdx.x >> 5] = lo[k]; s_hi[k][threadIdx.x >> 5] = hi[k]; }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        PairGeom g = geom[p];
        double ext[3];
        for (int k = 0; k < 3; ++k) {
            double l = fmin(fmin(s_lo[k][0], s_lo[k][1]), fmin(s_lo[k][2], s_lo[k][3]));
            double h = fmax(fmax(s_hi[k][0], s_hi[k][1]), fmax(s_hi[k][2], s_hi[k][3]));
            if (!(l <= h)) { l = 0.0; h = 0.0; }
            g.lo[k] = l; ext[k] = h - l;
        }
        const double e0 = fmax(ext[0], fmax(ext[1], ext[2]));
        // the clouds are voxel-lattice surfaces: ~3 voxels per cell puts the 20 nearest neighbours inside the
        // 27-cell neighbourhood without dragging hundreds of candidates along (a bounding-box estimate is
        // inflated by the very outliers the filter is there to find)
        double h = P.sor_cell;
        const double hmin = e0 / (double)((1 << kFbCellBits) - 1);
        if (h < hmin) h = hmin;
        if (!(h > 0.0) || !isfinite(h)) h = 1.0;
        for (int k = 0; k < 3; ++k) {
            int dim = (int)floor(ext[k] / h) + 1;
            g.dim[k] = dim < 1 ? 1 : (dim > (1 << kFbCellBits) ? (1 << kFbCellBits) : dim);
        }
        g.h = h; g.inv_h = 1.0 / h;
        g.thr = 0.0; g.cols = (g.dim[0] * g.dim[1] <= kFbColCap) ? 1 : 0;
        geom[p] = g;
    }
}

__device__ __forceinline__ void load_pair_geom(GridGeom &g, const PairGeom &pg, int p, const int *col_tab = nullptr) {
    for (int k = 0; k < 3; ++k) { g.lo[k] = pg.lo[k]; g.dim[k] = pg.dim[k]; g.bits[k] = kFbCellBits; }
    g.h = pg.h; g.inv_h = pg.inv_h;
    g.prefix = (unsigned long long)p << kFbCodeBits;
    g.cols = (col_tab != nullptr && pg.cols) ? col_tab + (size_t)p * (kFbColCap + 1) : nullptr;
}

// Column table of every group: one thread per sorted point; the first point of a column (and of the group) writes
// its index into the entries of all columns since the previous point's, the group's last point closes the table.
__global__ void __launch_bounds__(kFbThreads) fb_sor_cols_kernel(const unsigned long long *__restrict__ keys, const int *n_dev,
                                                                 const PairGeom *__restrict__ geom, int *__restrict__ col_tab) {
    const int n = *n_dev;
    const long long stride = (long long)gridDim.x * kFbThreads;
    for (long long j = (long long)blockIdx.x * kFbThreads + threadIdx.x; j < n; j += stride) {
        const unsigned long long key = keys[j];
        const int p = (int)(key >> kFbCodeBits);
        const PairGeom pg = geom[p];
        if (!pg.cols || pg.skip) continue;
        const int col = (int)((key >> (2 * kFbCellBits)) & ((1u << kFbCellBits) - 1u)) * pg.dim[1] +
                        (int)((key >> kFbCellBits) & ((1u << kFbCellBits) - 1u));
        int prev = -1;
        if (j > pg.start) {
            const unsigned long long kp = keys[j - 1];
            prev = (int)((kp >> (2 * kFbCellBits)) & ((1u << kFbCellBits) - 1u)) * pg.dim[1] + (int)((kp >> kFbCellBits) & ((1u << kFbCellBits) - 1u));
        }
        int *tab = col_tab + (size_t)p * (kFbColCap + 1);
        for (int c = prev + 1; c <= col; ++c) tab[c] = (int)j;
        if (j == (long long)pg.start + pg.n - 1)
            for (int c = col + 1; c <= pg.dim[0] * pg.dim[1]; ++c) tab[c] = (int)j + 1;
    }
}

__global__ void __launch_bounds__(kFbThreads) fb_sor_cellkey_kernel(const FbParams P, const double *__restrict__ pd, const int *__restrict__ grp,
                                                                    const int *n_dev, const PairGeom *__restrict__ geom,
                                                                    int npass, unsigned long long *keys, unsigned *vals,
                                                                    unsigned *hist) {
    __shared__ unsigned s_hist[kRsMaxPasses * 256];
    const int n = *n_dev;
    for (int i = threadIdx.x; i < kRsMaxPasses * 256; i += kFbThreads) s_hist[i] = 0;
    __syncthreads();
    const long long stride = (long long)gridDim.x * kFbThreads;
    for (long long base = (long long)blockIdx.x * kFbThreads; base < n; base += stride) {
        const long long i = base + threadIdx.x;
        const bool valid = i < n;
        unsigned long long key = 0;
        if (valid) {
            const int p = grp[i];
            GridGeom g;
            const PairGeom pg = geom[p];
            load_pair_geom(g, pg, p);
            // a group that is never filtered keeps cell 0: its points only have to stay inside the group's range
            key = pg.skip ? g.pack(0, 0, 0) : g.pack(g.cell(pd[3 * i], 0), g.cell(pd[3 * i + 1], 1), g.cell(pd[3 * i + 2], 2));
            keys[i] = key;
            vals[i] = (unsigned)i;
        }
        rs_hist_add(s_hist, key, valid, npass);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < npass * 256; i += kFbThreads) {
        const unsigned c = s_hist[i];
        if (c) atomicAdd(&hist[i], c);
    }
}

__global__ void fb_gather_f64_kernel(const double *__restrict__ pd, const int *n_dev, const unsigned *__restrict__ vals,
                                     double *__restrict__ ps) {
    const int n = *n_dev;
    for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < n; j += (long long)gridDim.x * blockDim.x) {
        const long long i = vals[j];
        ps[3 * j] = pd[3 * i]; ps[3 * j + 1] = pd[3 * i + 1]; ps[3 * j + 2] = pd[3 * i + 2];
    }
}

// ---- lattice neighbour search on occupancy bit planes
// The filter's inputs here are voxel-lattice points (V1 output), one or two camera clouds stacked, so a point's
// neighbours can be found by probing lattice offsets in shell order (dx^2 + dy^2 + dz^2 = 0, 1, 2, 3, 4, 5, 6, 8, 9:
// 123 offsets) instead of collecting and ranking every point of the surrounding grid cells.  Every group gets two
// dense bit planes over the core of its bounding box (mean +- 3 sigma per axis: the outliers the filter is there to
// find would otherwise inflate the box a hundredfold) -- plane A: voxel occupied, plane B: occupied by both clouds
// (vstack keeps duplicates, and a duplicate is a neighbour at distance 0).  A dozen kilobytes per object at 5 mm, so
// the probes stay in L1 / L2.  Shells are separated by far more than the fp32 rounding of the coordinates
// (tests/test_lattice_shells.py), so the shell holding the k-th neighbour follows from counts alone; only the <= 32
// members of the shells up to it get exact fp64 distances (from the voxel indices: a lattice point IS fl32(k * v)),
// are ranked and summed ascending like Open3D's std::accumulate.  What the planes cannot settle -- fewer than k
// neighbours within 3 voxels, more than 32 candidates, a neighbourhood reaching outside the core while points live
// there, groups whose planes do not fit the budget -- is flagged for the grid search (fb_sor_knn_kernel).
// offsets packed (dx + 4) | (dy + 4) << 4 | (dz + 4) << 8, sorted by shell
__device__ const unsigned short kLatOff[128] = {
    0x444, 0x443, 0x434, 0x344, 0x544, 0x454, 0x445, 0x433, 0x343, 0x543, 0x453, 0x334, 0x534, 0x354, 0x554, 0x435,
    0x345, 0x545, 0x455, 0x333, 0x533, 0x353, 0x553, 0x335, 0x535, 0x355, 0x555, 0x442, 0x424, 0x244, 0x644, 0x464,
    0x446, 0x432, 0x342, 0x542, 0x452, 0x423, 0x243, 0x643, 0x463, 0x324, 0x524, 0x234, 0x634, 0x254, 0x654, 0x364,
    0x564, 0x425, 0x245, 0x645, 0x465, 0x436, 0x346, 0x546, 0x456, 0x332, 0x532, 0x352, 0x552, 0x323, 0x523, 0x233,
    0x633, 0x253, 0x653, 0x363, 0x563, 0x325, 0x525, 0x235, 0x635, 0x255, 0x655, 0x365, 0x565, 0x336, 0x536, 0x356,
    0x556, 0x422, 0x242, 0x642, 0x462, 0x224, 0x624, 0x264, 0x664, 0x426, 0x246, 0x646, 0x466, 0x441, 0x322, 0x522,
    0x232, 0x632, 0x252, 0x652, 0x362, 0x562, 0x223, 0x623, 0x263, 0x663, 0x414, 0x144, 0x744, 0x474, 0x225, 0x625,
    0x265, 0x665, 0x326, 0x526, 0x236, 0x636, 0x256, 0x656, 0x366, 0x566, 0x447, 0, 0, 0, 0, 0};
// 123 offsets; the shells end after 1, 7, 19, 27, 33, 57, 81, 93, 123 of them
constexpr int kLatReach = 3;       // largest |offset| per axis

struct LatGeom {       // bit planes of one group
    int x0, y0, z0;    // first voxel (biased index) of the core box
    int nx, ny, nzw;   // extent in x, y; 32-bit words per z column
    unsigned off;      // first word of plane A; plane B follows nx * ny * nzw words later
    int enabled;       // planes exist
    int open;          // some point of the group lies outside the core box
};

// one block per group: core box = bounding box clipped to mean +- (3 sigma + 4) voxels per axis; plane allocation
__global__ void __launch_bounds__(128) fb_lat_plan_kernel(const FbParams P, const unsigned long long *__restrict__ gcode, const PairGeom *__restrict__ geom,
                                                          LatGeom *lat, unsigned *cursor, unsigned budget_words) {
    __shared__ double s_sum[6][4];
    __shared__ int s_min[3][4], s_max[3][4];
    const int g = blockIdx.x;
    const PairGeom pg = geom[g];
    LatGeom L;
    L.x0 = L.y0 = L.z0 = 0; L.nx = L.ny = L.nzw = 0; L.off = 0u; L.enabled = 0; L.open = 0;
    if (pg.n < 2 || pg.skip || pg.passthrough) { if (threadIdx.x == 0) lat[g] = L; return; }   // block-uniform
    const unsigned fm = (1u << P.obits) - 1u;
    double sum[6] = {0, 0, 0, 0, 0, 0};
    int mn[3] = {INT_MAX, INT_MAX, INT_MAX}, mx[3] = {INT_MIN, INT_MIN, INT_MIN};
    for (int i = pg.start + threadIdx.x; i < pg.start + pg.n; i += 128) {
        const unsigned long long c = gcode[i];
        const int k[3] = {(int)((c >> (2 * P.obits)) & fm), (int)((c >> P.obits) & fm), (int)(c & fm)};
#pragma unroll
        for (int a = 0; a < 3; ++a) { sum[a] += (double)k[a]; sum[3 + a] += (double)k[a] * (double)k[a]; mn[a] = min(mn[a], k[a]); mx[a] = max(mx[a], k[a]); }
    }
#pragma unroll
    for (int a = 0; a < 6; ++a) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum[a] += __shfl_xor_sync(kFullMask, sum[a], o);
        if (lane_id() == 0) s_sum[a][threadIdx.x >> 5] = sum[a];
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        mn[a] = __reduce_min_sync(kFullMask, mn[a]); mx[a] = __reduce_max_sync(kFullMask, mx[a]);
        if (lane_id() == 0) { s_min[a][threadIdx.x >> 5] = mn[a]; s_max[a][threadIdx.x >> 5] = mx[a]; }
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    int lo[3], hi[3];
    for (int a = 0; a < 3; ++a) {
        const double n = (double)pg.n;
        const double m = (s_sum[a][0] + s_sum[a][1] + s_sum[a][2] + s_sum[a][3]) / n;
        const double var = fmax((s_sum[3 + a][0] + s_sum[3 + a][1] + s_sum[3 + a][2] + s_sum[3 + a][3]) / n - m * m, 0.0);
        const int half = (int)ceil(3.0 * sqrt(var)) + 4;
        const int bmin = min(min(s_min[a][0], s_min[a][1]), min(s_min[a][2], s_min[a][3]));
        const int bmax = max(max(s_max[a][0], s_max[a][1]), max(s_max[a][2], s_max[a][3]));
        lo[a] = max(bmin, (int)floor(m) - half);
        hi[a] = min(bmax, (int)ceil(m) + half);
        if (lo[a] > bmin || hi[a] < bmax) L.open = 1;
    }
    L.x0 = lo[0]; L.y0 = lo[1]; L.z0 = lo[2];
    L.nx = hi[0] - lo[0] + 1; L.ny = hi[1] - lo[1] + 1; L.nzw = (hi[2] - lo[2] + 32) >> 5;
    const unsigned long long words = (unsigned long long)L.nx * L.ny * L.nzw;
    if (2ull * words <= (unsigned long long)budget_words) {
        const unsigned at = atomicAdd(cursor, (unsigned)(2ull * words));
        if ((unsigned long long)at + 2ull * words <= (unsigned long long)budget_words) { L.off = at; L.enabled = 1; }
    }
    lat[g] = L;
}

// the words the plan kernel handed out, and only those, start from zero
__global__ void __launch_bounds__(kFbThreads) fb_lat_zero_kernel(const unsigned *cursor, unsigned budget_words, uint4 *bits) {
    const unsigned used = min(*cursor, budget_words);
    const long long n4 = ((long long)used + 3) >> 2, stride = (long long)gridDim.x * kFbThreads;
    for (long long i = (long long)blockIdx.x * kFbThreads + threadIdx.x; i < n4; i += stride) bits[i] = make_uint4(0u, 0u, 0u, 0u);
}

// one thread per point: its bit in plane A; the second point of a voxel (the other camera's) sets plane B
__global__ void __launch_bounds__(kFbThreads) fb_lat_fill_kernel(const FbParams P, const unsigned long long *__restrict__ gcode, const int *__restrict__ grp,
                                                                 const int *n_dev, const LatGeom *__restrict__ lat, unsigned *bits) {
    const int n = *n_dev;
    const unsigned fm = (1u << P.obits) - 1u;
    const long long stride = (long long)gridDim.x * kFbThreads;
    for (long long i = (long long)blockIdx.x * kFbThreads + threadIdx.x; i < n; i += stride) {
        const LatGeom L = lat[grp[i]];
        if (!L.enabled) continue;
        const unsigned long long c = gcode[i];
        const int x = (int)((c >> (2 * P.obits)) & fm) - L.x0, y = (int)((c >> P.obits) & fm) - L.y0, z = (int)(c & fm) - L.z0;
        if ((unsigned)x >= (unsigned)L.nx || (unsigned)y >= (unsigned)L.ny || (unsigned)z >= (unsigned)(L.nzw << 5)) continue;
        const unsigned idx = L.off + (unsigned)((x * L.ny + y) * L.nzw + (z >> 5)), bit = 1u << (z & 31);
        if (atomicOr(&bits[idx], bit) & bit) atomicOr(&bits[idx + (unsigned)(L.nx * L.ny * L.nzw)], bit);
    }
}

// One THREAD per query (queries in cell-sorted order, so neighbouring threads probe the same bit-plane words): count the
// occupied offsets shell by shell until k neighbours are reached, then give the members of those shells their exact
// distances, order them (only members of one shell can be out of order) and sum the square roots ascending.  A warp per
// query spends ~500 warp instructions on scans, a 32 x 32 ranking and shuffles; a thread per query needs ~2-3 thousand
// of its own, i.e. under a hundred warp-instruction equivalents.
__constant__ unsigned short c_lat_off[128];   // copy of kLatOff (uniform index across a warp -> constant-cache broadcast)
constexpr int kLatThreads = 128;
__global__ void __launch_bounds__(kLatThreads) fb_sor_lattice_kernel(const FbParams P, const unsigned long long *__restrict__ gcode, const int *__restrict__ grp,
                                                                     const unsigned *__restrict__ svals, const int *n_dev, const PairGeom *__restrict__ geom,
                                                                     const LatGeom *__restrict__ lat, const unsigned *__restrict__ bits,
                                                                     double *__restrict__ avg, unsigned char *need, int *n_settled) {
    const int n = *n_dev;
    const int fm = (1 << P.obits) - 1;
    int settled = 0;
    const long long stride = (long long)gridDim.x * kLatThreads;
    for (long long j = (long long)blockIdx.x * kLatThreads + threadIdx.x; j < n; j += stride) {
        const int i = (int)svals[j];
        const int g = grp[i];
        const LatGeom L = lat[g];
        need[j] = 1;                                    // until settled below: grid search (a never-filtered group is skipped there)
        if (!L.enabled) continue;
        const unsigned long long code = gcode[i];
        const int x = (int)((code >> (2 * P.obits)) & (unsigned long long)fm), y = (int)((code >> P.obits) & (unsigned long long)fm),
                  z = (int)(code & (unsigned long long)fm);
        const int lx = x - L.x0, ly = y - L.y0, lz = z - L.z0, nz = L.nzw << 5;
        // with points outside the core box, only a neighbourhood that lies wholly inside it can be trusted
        const bool inside = lx >= kLatReach && ly >= kLatReach && lz >= kLatReach && lx < L.nx - kLatReach && ly < L.ny - kLatReach && lz < nz - kLatReach;
        if (L.open && !inside) continue;
        const PairGeom pg = geom[g];
        const int kk = min(P.sor_k, pg.n);
        const bool dup = pg.n0 != pg.n;                  // two clouds stacked: a voxel may hold two points (plane B)
        const unsigned plane = (unsigned)(L.nx * L.ny * L.nzw);
        // ---- pass 1: occupancy of the offsets, shell by shell, until kk neighbours are in
        unsigned ha[4] = {0u, 0u, 0u, 0u}, hb[4] = {0u, 0u, 0u, 0u};   // offsets occupied once / twice
        int cum = 0, last = -1;
        constexpr int kShellEnd[9] = {0, 6, 18, 26, 32, 56, 80, 92, 122};
        int oi = 0;
#pragma unroll 1
        for (int sh = 0; sh < 9 && last < 0; ++sh) {
#pragma unroll 1
            for (; oi <= kShellEnd[sh]; ++oi) {
                const unsigned o = c_lat_off[oi];
                const int ax = lx + (int)(o & 15u) - 4, ay = ly + (int)((o >> 4) & 15u) - 4, az = lz + (int)((o >> 8) & 15u) - 4;
                if ((unsigned)ax < (unsigned)L.nx && (unsigned)ay < (unsigned)L.ny && (unsigned)az < (unsigned)nz) {
                    const unsigned idx = L.off + (unsigned)((ax * L.ny + ay) * L.nzw + (az >> 5));
                    if ((bits[idx] >> (az & 31)) & 1u) {
                        ha[oi >> 5] |= 1u << (oi & 31);
                        ++cum;
                        if (dup && ((bits[idx + plane] >> (az & 31)) & 1u)) { hb[oi >> 5] |= 1u << (oi & 31); ++cum; }
                    }
                }
            }
            if (cum >= kk) last = sh;
        }
        if (last < 0 || cum > 32) {                      // not settled here; what the count says about the block the grid search needs:
            // fewer than kk within 3 voxels -> the 27-cell block (reach 3 voxels) cannot do; fewer than kk / 4 -> by the r^2
            // growth of a surface's neighbourhood the 125-cell block (reach 6) will not either.  (A hint: any block that
            // reaches the kk-th neighbour gives the exact answer.)
            if (last < 0) need[j] = 4 * cum < kk ? 3 : 2;
            continue;
        }
        // ---- pass 2: exact fp64 distances of the members (a lattice point is (double) fl32(fl32(k) * v)), kept ascending
        double dl[32];
        int m = 0;
        const int kx = x - P.obias, ky = y - P.obias, kz = z - P.obias;
        const double qx = (double)__fmul_rn((float)kx, P.v32), qy = (double)__fmul_rn((float)ky, P.v32), qz = (double)__fmul_rn((float)kz, P.v32);
#pragma unroll 1
        for (int w = 0; w < 4; ++w) {
#pragma unroll 1
            for (unsigned bitsw = ha[w]; bitsw; bitsw &= bitsw - 1) {
                const int b = __ffs((int)bitsw) - 1;
                const unsigned o = c_lat_off[32 * w + b];
                const double p[3] = {(double)__fmul_rn((float)(kx + (int)(o & 15u) - 4), P.v32), (double)__fmul_rn((float)(ky + (int)((o >> 4) & 15u) - 4), P.v32),
                                     (double)__fmul_rn((float)(kz + (int)((o >> 8) & 15u) - 4), P.v32)};
                const double d2 = dist2_nanoflann(qx, qy, qz, p);
                const int copies = 1 + (int)((hb[w] >> b) & 1u);
                for (int c = 0; c < copies; ++c) {       // insertion from the back: members arrive shell by shell
                    int t = m++;
                    while (t > 0 && dl[t - 1] > d2) { dl[t] = dl[t - 1]; --t; }
                    dl[t] = d2;
                }
            }
        }
        double sum = 0.0;
        for (int t = 0; t < kk; ++t) sum = __dadd_rn(sum, sqrt(dl[t]));   // ascending, from 0.0, like Open3D
        avg[i] = __ddiv_rn(sum, (double)kk);
        need[j] = 0;
        ++settled;
    }
    settled = __reduce_add_sync(__activemask(), settled);
    if (settled && lane_id() == __ffs(__activemask()) - 1) atomicAdd(n_settled, settled);
}

// One warp per query; queries are handed out by an atomic ticket (their cost varies by two orders of magnitude
// between a surface point and an isolated outlier), the next ticket being requested while the current query runs.
#ifndef RT3D_KNN_MASKED_BLOCKS
#define RT3D_KNN_MASKED_BLOCKS 4   // blocks per SM of the masked variant (tuning experiments override it)
#endif
template <int SLOTS, bool MASKED>
__global__ void __launch_bounds__(kFbThreads, MASKED ? RT3D_KNN_MASKED_BLOCKS : 6) fb_sor_knn_kernel(const double *__restrict__ ps, const unsigned long long *__restrict__ keys,
                                                                const unsigned *__restrict__ vals, const int *n_dev, int k,
                                                                const PairGeom *__restrict__ geom, const int *__restrict__ col_tab,
                                                                double *__restrict__ avg, int *ticket, const unsigned char *__restrict__ need) {
    __shared__ KnnWarpScratch s_knn[kFbThreads / 32];
    const int n = *n_dev;
    const int lane = lane_id();
    KnnCache cache = {~0ull, 0, 0};
    // guided self-scheduling: the ticket counts queries; a warp takes up to kFbKnnChunk consecutive ones while plenty
    // remain (neighbours in cell order share their candidate list) and single queries towards the end
    const int warps = (int)((gridDim.x * (unsigned)kFbThreads) >> 5);
    int take = min(kFbKnnChunk, max(1, n / (4 * warps)));
    int tk = 0;
    if (lane == 0) tk = atomicAdd(ticket, take);
    for (;;) {
        const int j0 = __shfl_sync(kFullMask, tk, 0);
        if (j0 >= n) break;
        const int j1 = min(j0 + take, n);
        take = min(kFbKnnChunk, max(1, (n - j0) / (4 * warps)));
        if (lane == 0) tk = atomicAdd(ticket, take);
        for (int j = j0; j < j1; ++j) {
            const int hint = MASKED ? need[j] : 1;
            if (MASKED && hint == 0) continue;   // settled on the bit planes
            const unsigned long long key = keys[j];
            const int p = (int)(key >> kFbCodeBits);
            const PairGeom pg = geom[p];
            if (pg.skip | pg.passthrough) continue;   // never filtered: nobody reads its mean distances
            GridGeom g;
            load_pair_geom(g, pg, p, col_tab);
            int cx, cy, cz;
            g.unpack(key, cx, cy, cz);
            const int kk = k < pg.n ? k : pg.n;
            // groups are contiguous and in the same order before and after the sort (group id = high key bits)
            const double m = knn_mean_distance<SLOTS>(g, keys, ps, pg.start, pg.start + pg.n, ps[3 * (size_t)j], ps[3 * (size_t)j + 1],
                                                      ps[3 * (size_t)j + 2], cx, cy, cz, kk, nullptr, &s_knn[threadIdx.x >> 5], &cache, key,
                                                      hint == 3 ? 4 : (hint == 2 ? 2 : 1));
            if (lane == 0) avg[vals[j]] = m;
        }
    }
}

// one block per group: Open3D's sequential mean / sigma over the group's points in input order.  The sums
// must run in index order to round like std::accumulate, so one thread carries the chain; the block only
// stages the values in shared memory so that the chain is bound by the fp64 add latency alone.
constexpr int kStatChunk = 2048;
__global__ void __launch_bounds__(kFbThreads) fb_sor_stats_kernel(const double *__restrict__ avg, double std_ratio, PairGeom *geom) {
    __shared__ double s_buf[kStatChunk];
    __shared__ double s_mu;
    const int p = blockIdx.x;
    const long long a = geom[p].start, n = geom[p].n;
    if (n == 0 || geom[p].skip || geom[p].passthrough) return;
    double acc = 0.0;
    for (int pass = 0; pass < 2; ++pass) {
        acc = 0.0;
        const double mu = pass ? s_mu : 0.0;
        for (long long base = 0; base < n; base += kStatChunk) {
            const int m = (int)min((long long)kStatChunk, n - base);
            for (int i = threadIdx.x; i < m; i += kFbThreads) {
                const double x = avg[a + base + i];
                double term = 0.0;
                if (x > 0.0) { const double dlt = __dsub_rn(x, mu); term = pass ? __dmul_rn(dlt, dlt) : x; }
                s_buf[i] = term;      // "+ 0.0" leaves the running sum unchanged, like the skipped element
            }
            __syncthreads();
            if (threadIdx.x == 0) {
#pragma unroll 8
                for (int i = 0; i < m; ++i) acc = __dadd_rn(acc, s_buf[i]);
            }
            __syncthreads();
        }
        if (threadIdx.x == 0 && pass == 0) s_mu = __ddiv_rn(acc, (double)n);
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const double sigma = sqrt(__ddiv_rn(acc, __dsub_rn((double)n, 1.0)));
        geom[p].thr = __dadd_rn(s_mu, __dmul_rn(std_ratio, sigma));
    }
}

// survivors of every group, compacted in group order (= the order of the group-ordered arrays)
struct FbSorSelectOp {
    using Payload = int;
    const double *avg, *pd;
    const int *grp;
    const PairGeom *geom;
    double *fused64;     // survivors as fp64 rows ...
    float *fused32;      // ... and as fp32 rows for the subtraction (RT3D_FUSE_BY_INDEX: these are the outputs) or nullptr
    int *fused_grp;
    __device__ __forceinline__ unsigned eval4(long long i0, long long n, int (&p)[4]) const {
        unsigned keep = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const long long i = i0 + j;
            if (i >= n) break;
            const int g = grp[i];
            const double a = avg[i];
            if (geom[g].passthrough || (!geom[g].skip && a > 0.0 && a < geom[g].thr)) { p[j] = (int)i; keep |= 1u << j; }
        }
        return keep;
    }
    __device__ __forceinline__ void store(long long pos, const int &i) const {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const double v = pd[3 * (long long)i + k];
            fused64[3 * pos + k] = v;
            if (fused32) fused32[3 * pos + k] = (float)v;   // exact: the values originate from fp32 lattice points
        }
        fused_grp[pos] = grp[i];
    }
};

// RT3D_FUSE_BY_INDEX: the groups are the outputs; only the optional meta rows remain to be written
__global__ void fb_meta_index_kernel(const FbParams P, const PairGeom *__restrict__ geom, const int *__restrict__ obj_off, int *obj_meta) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= P.F * P.slots) return;
    const int n = obj_off[g + 1] - obj_off[g];
    obj_meta[4 * g] = geom[g].kind == kGrpEmpty ? RT3D_CLOUD_EMPTY : (geom[g].passthrough ? RT3D_CLOUD_LEFTOVER : RT3D_CLOUD_PAIR);
    obj_meta[4 * g + 1] = geom[g].cls; obj_meta[4 * g + 2] = n; obj_meta[4 * g + 3] = 0;
}

// ---- D.2 RT3D_FUSE_CENTROID: centroids.  calculate_centroid = np.mean(axis=0) (:74-75): a sequential row sum in the
// array's dtype, then / n.  Camera-1 clouds are measured AFTER the filter (fp64, :131-136), camera-2 clouds
// unfiltered (fp32, :139-140).  One block per group; three threads carry the x / y / z chains over rows the block
// stages in shared memory.  An empty filtered cloud gives 0 / 0 = NaN, like NumPy: it matches nothing.
constexpr int kCenChunk = 512;
__global__ void __launch_bounds__(128) fb_fuse_centroid_kernel(const PairGeom *__restrict__ geom, const int *__restrict__ filt_off,
                                                               const double *__restrict__ filt64, const int *__restrict__ seg_off,
                                                               const float *__restrict__ obj_pts, double *__restrict__ cen) {
    __shared__ double s_buf[3 * kCenChunk];
    const int g = blockIdx.x;
    const int kind = geom[g].kind;
    if ((kind != kGrpCam1 && kind != kGrpCam2) || geom[g].skip) return;
    const bool f32 = kind == kGrpCam2;
    const long long a = f32 ? seg_off[geom[g].seg] : filt_off[g];
    const long long n = f32 ? geom[g].n : filt_off[g + 1] - filt_off[g];
    double acc = 0.0;
    float accf = 0.0f;
    for (long long base = 0; base < n; base += kCenChunk) {
        const int m = (int)min((long long)kCenChunk, n - base);
        for (int i = threadIdx.x; i < 3 * m; i += 128)
            s_buf[i] = f32 ? (double)obj_pts[3 * (a + base) + i] : filt64[3 * (a + base) + i];
        __syncthreads();
        if (threadIdx.x < 3) {
            if (f32) {
#pragma unroll 8
                for (int i = 0; i < m; ++i) accf = __fadd_rn(accf, (float)s_buf[3 * i + threadIdx.x]);
            } else {
#pragma unroll 8
                for (int i = 0; i < m; ++i) acc = __dadd_rn(acc, s_buf[3 * i + threadIdx.x]);
            }
        }
        __syncthreads();
    }
    if (threadIdx.x < 3) cen[3 * g + threadIdx.x] = f32 ? (double)__fdiv_rn(accf, (float)n) : __ddiv_rn(acc, (double)n);
}

// ---- D.3 RT3D_FUSE_CENTROID: the greedy per-class matching (:117-163), one thread per frame (<= 2 n_obj clouds).
// Output slot = one entry of the reference's fused_point_clouds list, as up to two row ranges.
struct FbOutCloud {
    int a_start, a_n;   // first part: rows of filt64 (kind PAIR / MATCHED / UNMATCHED) or of obj_pts (LEFTOVER)
    int b_start, b_n;   // second part (MATCHED): the filtered match, rows of filt64
    int kind, cls;
};

// np.array_equal(pc2, best_match) (:62-63, :155): pc2 = unfiltered camera-2 cloud (fp32 rows), best_match = filtered copy
__device__ bool fb_cloud_equals(const float *__restrict__ u, const double *__restrict__ f, int n) {
    for (int i = 0; i < 3 * n; ++i)
        if ((double)u[i] != f[i]) return false;
    return true;
}

__global__ void __launch_bounds__(64) fb_fuse_match_kernel(const FbParams P, const PairGeom *__restrict__ geom, const int *__restrict__ filt_off,
                                                           const double *__restrict__ filt64, const int *__restrict__ seg_off,
                                                           const float *__restrict__ obj_pts, const double *__restrict__ cen,
                                                           FbOutCloud *out, int *slot_n) {
    const int f = blockIdx.x * 64 + threadIdx.x;
    if (f >= P.F) return;
    const int g0 = f * P.slots, g1 = g0 + P.slots;
    int slot = g0;
    auto emit = [&](int kind, int cls, int as, int an, int bs, int bn) {
        FbOutCloud o;
        o.a_start = as; o.a_n = an; o.b_start = bs; o.b_n = bn; o.kind = kind; o.cls = cls;
        out[slot] = o; slot_n[slot] = an + bn;
        ++slot;
    };
    int g = g0;
    while (g < g1 && geom[g].kind != kGrpEmpty) {
        const int cls = geom[g].cls;
        if (geom[g].kind == kGrpPair) {
            emit(RT3D_CLOUD_PAIR, cls, filt_off[g], filt_off[g + 1] - filt_off[g], 0, 0);
            ++g;
            continue;
        }
        int c1b = g, c1e = g;
        while (c1e < g1 && geom[c1e].kind == kGrpCam1 && geom[c1e].cls == cls) ++c1e;
        int c2e = c1e;
        while (c2e < g1 && geom[c2e].kind == kGrpCam2 && geom[c2e].cls == cls) ++c2e;
        unsigned long long alive = (c2e - c1e) >= 64 ? ~0ull : ((1ull << (c2e - c1e)) - 1ull);   // camera-2 clouds still in pcs2
        for (int a = c1b; a < c1e; ++a) {
            const double x1 = cen[3 * a], y1 = cen[3 * a + 1], z1 = cen[3 * a + 2];
            double best = __longlong_as_double(0x7ff0000000000000ll);   // float('inf')
            int bm = -1;
            for (int b = c1e; b < c2e; ++b) {
                if (!((alive >> (b - c1e)) & 1ull)) continue;
                // np.linalg.norm(c1 - c2) = sqrt(dot(d, d)); the 3-element dot product accumulates with FMAs
                // (measured against NumPy 2.3 / OpenBLAS in this image: x0*x0, then fma, fma)
                const double dx = __dsub_rn(x1, cen[3 * b]), dy = __dsub_rn(y1, cen[3 * b + 1]), dz = __dsub_rn(z1, cen[3 * b + 2]);
                const double d = sqrt(__fma_rn(dz, dz, __fma_rn(dy, dy, __dmul_rn(dx, dx))));
                if (d < best && d < P.fuse_thr) { best = d; bm = b; }
            }
            const int an = filt_off[a + 1] - filt_off[a];
            if (bm < 0) { emit(RT3D_CLOUD_UNMATCHED, cls, filt_off[a], an, 0, 0); continue; }
            const int bn = filt_off[bm + 1] - filt_off[bm];
            emit(RT3D_CLOUD_MATCHED, cls, filt_off[a], an, filt_off[bm], bn);
            // pcs2 = [pc for pc in pcs2 if not point_clouds_equal(pc, best_match)] (:155)
            for (int b = c1e; b < c2e; ++b) {
                if (!((alive >> (b - c1e)) & 1ull) || geom[b].n != bn) continue;
                if (b == bm || fb_cloud_equals(obj_pts + 3 * (size_t)seg_off[geom[b].seg], filt64 + 3 * (size_t)filt_off[bm], bn))
                    alive &= ~(1ull << (b - c1e));
            }
        }
        for (int b = c1e; b < c2e; ++b)
            if ((alive >> (b - c1e)) & 1ull) emit(RT3D_CLOUD_LEFTOVER, cls, seg_off[geom[b].seg], geom[b].n, 0, 0);
        g = c2e;
    }
    for (; slot < g1; ++slot) {
        FbOutCloud o;
        o.a_start = o.a_n = o.b_start = o.b_n = 0; o.kind = RT3D_CLOUD_EMPTY; o.cls = -1;
        out[slot] = o; slot_n[slot] = 0;
    }
}

// one block: obj_off = exclusive scan of the slot sizes; the total is clamped to the capacity of obj_xyz
__global__ void __launch_bounds__(kFbThreads) fb_fuse_scan_kernel(const int *__restrict__ slot_n, int nslots, int cap, int *obj_off, FbCounts *cnt) {
    __shared__ int s_tmp[8];
    __shared__ int s_run;
    if (threadIdx.x == 0) s_run = 0;
    __syncthreads();
    for (int base = 0; base < nslots; base += kFbThreads) {
        const int i = base + threadIdx.x;
        const int v = i < nslots ? slot_n[i] : 0;
        int total;
        const int excl = block_excl_scan_256(v, s_tmp, total);
        const int run = s_run;
        if (i < nslots) obj_off[i] = min(run + excl, cap);
        __syncthreads();
        if (threadIdx.x == 0) s_run = min(run + total, cap + 1);
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        if (s_run > cap) { atomicOr(&cnt->status, RT3D_DEV_CAPACITY); s_run = cap; }
        obj_off[nslots] = s_run; cnt->fused = s_run;
    }
}

// the fused clouds, row by row: slot by binary search in obj_off, then part a / part b of the slot
__global__ void __launch_bounds__(kFbThreads) fb_fuse_assemble_kernel(const FbOutCloud *__restrict__ out, const int *__restrict__ obj_off, int nslots,
                                                                      const double *__restrict__ filt64, const float *__restrict__ obj_pts,
                                                                      double *__restrict__ obj_xyz, float *__restrict__ fused32, int *obj_meta) {
    const int total = obj_off[nslots];
    const long long stride = (long long)gridDim.x * kFbThreads;
    for (long long i = (long long)blockIdx.x * kFbThreads + threadIdx.x; i < total; i += stride) {
        int lo = 0, hi = nslots;   // largest slot with obj_off[slot] <= i
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (obj_off[mid] <= (int)i) lo = mid; else hi = mid; }
        const FbOutCloud o = out[lo];
        const int r = (int)i - obj_off[lo];
        double x, y, z;
        if (o.kind == RT3D_CLOUD_LEFTOVER) {
            const float *s = obj_pts + 3 * (size_t)(o.a_start + r);
            x = (double)s[0]; y = (double)s[1]; z = (double)s[2];
        } else {
            const double *s = filt64 + 3 * (size_t)(r < o.a_n ? o.a_start + r : o.b_start + (r - o.a_n));
            x = s[0]; y = s[1]; z = s[2];
        }
        obj_xyz[3 * i] = x; obj_xyz[3 * i + 1] = y; obj_xyz[3 * i + 2] = z;
        fused32[3 * i] = (float)x; fused32[3 * i + 1] = (float)y; fused32[3 * i + 2] = (float)z;
    }
    if (obj_meta != nullptr)
        for (long long s = (long long)blockIdx.x * kFbThreads + threadIdx.x; s < nslots; s += stride) {
            const FbOutCloud o = out[s];
            obj_meta[4 * s] = o.kind; obj_meta[4 * s + 1] = o.cls; obj_meta[4 * s + 2] = o.a_n; obj_meta[4 * s + 3] = o.b_n;
        }
}

// masks handed over as the ultralytics fp32 planes (:295): non-zero (NaN included, like torch.nonzero) -> 1
__global__ void __launch_bounds__(kFbThreads) fb_mask_f32_kernel(const float4 *__restrict__ in, long long n4, unsigned *__restrict__ out) {
    const long long stride = (long long)gridDim.x * kFbThreads;
    for (long long i = (long long)blockIdx.x * kFbThreads + threadIdx.x; i < n4; i += stride) {
        const float4 m = ldg_stream_f4(reinterpret_cast<const float *>(in + i));
        out[i] = (m.x != 0.0f ? 1u : 0u) | (m.y != 0.0f ? 0x100u : 0u) | (m.z != 0.0f ? 0x10000u : 0u) | (m.w != 0.0f ? 0x1000000u : 0u);
    }
}

// --------------------------------------------------------------------------------- E: subtract
// (object grid + per-row flags: launch_subtract in k4_fuse_subtract.cu)
struct FbKeepOp {
    struct Payload { float x, y, z; };
    const float *ws;
    const unsigned char *keep;
    float *kept_xyz;
    __device__ __forceinline__ unsigned eval4(long long i0, long long n, Payload (&p)[4]) const {
        unsigned k = 0;
        if (i0 + 4 <= n && ((i0 & 3) == 0) && ((reinterpret_cast<uintptr_t>(ws) | reinterpret_cast<uintptr_t>(keep)) & 15) == 0) {
            // four flags in one word, the four rows in three 16-byte loads (skipped when none is kept)
            const unsigned kb = *reinterpret_cast<const unsigned *>(keep + i0);
            if (kb == 0u) return 0u;
            const float4 *src = reinterpret_cast<const float4 *>(ws + 3 * i0);
            const float4 a = src[0], b = src[1], c = src[2];
            p[0].x = a.x; p[0].y = a.y; p[0].z = a.z;
            p[1].x = a.w; p[1].y = b.x; p[1].z = b.y;
            p[2].x = b.z; p[2].y = b.w; p[2].z = c.x;
            p[3].x = c.y; p[3].y = c.z; p[3].z = c.w;
            return ((kb & 0xffu) ? 1u : 0u) | ((kb & 0xff00u) ? 2u : 0u) | ((kb & 0xff0000u) ? 4u : 0u) | ((kb & 0xff000000u) ? 8u : 0u);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const long long i = i0 + j;
            if (i >= n) break;
            if (keep[i]) { p[j].x = ws[3 * i]; p[j].y = ws[3 * i + 1]; p[j].z = ws[3 * i + 2]; k |= 1u << j; }
        }
        return k;
    }
    __device__ __forceinline__ void store(long long pos, const Payload &p) const {
        float *o = kept_xyz + 3 * pos;
        o[0] = p.x; o[1] = p.y; o[2] = p.z;
    }
};

// Compaction of the kept rows in two sweeps instead of one decoupled look-back pass: the predicate is a byte
// already in memory, so counting it per tile first (42 MB) is cheaper than stalling whole blocks on the look-back
// (measured: 31 of 56 stalled warp-cycles per issue at the block barrier, 0.39 ms for 0.58 GB).
constexpr int kKeepRows = 2;                                   // rows of 128 items per warp and tile
constexpr int kKeepTile = kFbThreads * 4 * kKeepRows;          // 2048 workspace rows per block
constexpr int kKeepTileShift = 11;
static_assert((1 << kKeepTileShift) == kKeepTile, "tile shift");
constexpr int kKeepScanThreads = 1024;

__global__ void __launch_bounds__(kFbThreads) fb_keep_count_kernel(const unsigned char *__restrict__ keep, const int *n_dev, int *tile_cnt) {
    __shared__ int s_part[kFbThreads / 32];
    const int n = *n_dev;
    const long long t0 = (long long)blockIdx.x * kKeepTile;
    int c = 0;
    if (t0 < n) {
#pragma unroll
        for (int r = 0; r < kKeepRows; ++r) {
            const long long i0 = t0 + (long long)(threadIdx.x >> 5) * (128 * kKeepRows) + r * 128 + (threadIdx.x & 31) * 4;
            if (i0 + 4 <= n) {
                const unsigned kb = *reinterpret_cast<const unsigned *>(keep + i0);
                c += ((kb & 0xffu) ? 1 : 0) + ((kb & 0xff00u) ? 1 : 0) + ((kb & 0xff0000u) ? 1 : 0) + ((kb & 0xff000000u) ? 1 : 0);
            } else {
                for (long long i = i0; i < n; ++i) c += keep[i] ? 1 : 0;
            }
        }
    }
    c = __reduce_add_sync(kFullMask, c);
    if (lane_id() == 0) s_part[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < kFbThreads / 32; ++w) t += s_part[w];
        tile_cnt[blockIdx.x] = t;
    }
}

// one block: exclusive scan of the tile counts (in place: tile_cnt -> tile offset) and the total; 8 tiles per thread
constexpr int kKeepScanPer = 8;
__global__ void __launch_bounds__(kKeepScanThreads) fb_keep_scan_kernel(int *tile_cnt, const int *n_dev, int *total) {
    __shared__ int s_w[kKeepScanThreads / 32];
    __shared__ int s_run;
    const int ntiles = (int)(((long long)*n_dev + kKeepTile - 1) / kKeepTile);
    if (threadIdx.x == 0) s_run = 0;
    __syncthreads();
    for (int base = 0; base < ntiles; base += kKeepScanThreads * kKeepScanPer) {
        const int i0 = base + threadIdx.x * kKeepScanPer;
        int v[kKeepScanPer], mine = 0;
#pragma unroll
        for (int j = 0; j < kKeepScanPer; ++j) { v[j] = i0 + j < ntiles ? tile_cnt[i0 + j] : 0; mine += v[j]; }
        const int incl = warp_incl_scan(mine);
        if (lane_id() == 31) s_w[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x < 32) {
            const int w = s_w[threadIdx.x];
            const int wi = warp_incl_scan(w);
            s_w[threadIdx.x] = wi - w;
        }
        __syncthreads();
        const int run = s_run;
        int at = run + s_w[threadIdx.x >> 5] + incl - mine;
#pragma unroll
        for (int j = 0; j < kKeepScanPer; ++j) { if (i0 + j < ntiles) tile_cnt[i0 + j] = at; at += v[j]; }
        __syncthreads();
        if (threadIdx.x == kKeepScanThreads - 1) s_run = at;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = s_run;
}

__global__ void __launch_bounds__(kFbThreads) fb_keep_write_kernel(const FbKeepOp op, const int *n_dev, const int *__restrict__ tile_off) {
    __shared__ int s_wtot[kFbThreads / 32];
    __shared__ float s_out[3 * kKeepTile];
    const int n = *n_dev;
    const long long t0 = (long long)blockIdx.x * kKeepTile;
    if (t0 >= n) return;
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    FbKeepOp::Payload pay[kKeepRows][4];
    unsigned keep[kKeepRows];
    int row_excl[kKeepRows];
    int warp_total = 0;
#pragma unroll
    for (int r = 0; r < kKeepRows; ++r) {
        const long long i0 = t0 + (long long)warp * (128 * kKeepRows) + r * 128 + lane * 4;
        keep[r] = (i0 < n) ? op.eval4(i0, n, pay[r]) : 0u;
    }
#pragma unroll
    for (int r = 0; r < kKeepRows; ++r) {
        const int cnt = __popc(keep[r]);
        const int incl = warp_incl_scan(cnt);
        row_excl[r] = warp_total + incl - cnt;
        warp_total += __shfl_sync(kFullMask, incl, 31);
    }
    if (lane == 0) s_wtot[warp] = warp_total;
    __syncthreads();
    // the tile's kept rows are contiguous in the output: gather them in shared memory, then store them coalesced
    // (12-byte rows stored lane by lane cost one 32-byte sector request per 4 bytes)
    int local = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kFbThreads / 32; ++w) { const int c = s_wtot[w]; local += w < warp ? c : 0; total += c; }
#pragma unroll
    for (int r = 0; r < kKeepRows; ++r) {
        int pos = local + row_excl[r];
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if ((keep[r] >> j) & 1u) { float *o = s_out + 3 * pos++; o[0] = pay[r][j].x; o[1] = pay[r][j].y; o[2] = pay[r][j].z; }
    }
    __syncthreads();
    float *dst = op.kept_xyz + 3 * (long long)tile_off[blockIdx.x];
    for (int i = threadIdx.x; i < 3 * total; i += kFbThreads) dst[i] = s_out[i];
}

// kept rows per frame (one block per frame), then kept_off = exclusive scan (one small block)
__global__ void __launch_bounds__(kFbThreads) fb_frame_count_kernel(const unsigned char *__restrict__ keep, const int *__restrict__ ws_off,
                                                                    int C, int *frame_cnt) {
    __shared__ int s_part[8];
    const int f = blockIdx.x;
    const int a = ws_off[f * C], b = ws_off[(f + 1) * C];
    int c = 0;
    for (int i = a + threadIdx.x; i < b; i += kFbThreads) c += keep[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(kFullMask, c, o);
    if (lane_id() == 0) s_part[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) { int t = 0; for (int w = 0; w < 8; ++w) t += s_part[w]; frame_cnt[f] = t; }
}
__global__ void fb_frame_scan_kernel(const int *frame_cnt, int F, int *kept_off) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        int run = 0;
        for (int f = 0; f < F; ++f) { kept_off[f] = run; run += frame_cnt[f]; }
        kept_off[F] = run;
    }
}

// --------------------------------------------------------------------------------- B': workspace outlier filter (optional)
// 2cam/2cams_mask_cpu.py:530 filters the fused workspace (the cameras' voxel sets concatenated, :525) with
// remove_statistical_outlier(20, 1.5) before anything is subtracted from it.  Here: one outlier-filter group per frame,
// the same grid search as for the objects (stage D), survivors compacted into a scratch copy that the subtraction
// then reads instead of ws_xyz.
__device__ __forceinline__ int fb_frame_of_ws_row(const int *__restrict__ ws_off, int C, int F, int i) {
    int lo = 0, hi = F;   // largest f with ws_off[f * C] <= i
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (ws_off[mid * C] <= i) lo = mid; else hi = mid; }
    return lo;
}
__global__ void fb_ws_group_kernel(const FbParams P, const int *__restrict__ ws_off, PairGeom *geom) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= P.F) return;
    PairGeom e;
    e.lo[0] = e.lo[1] = e.lo[2] = 0.0; e.h = 1.0; e.inv_h = 1.0; e.dim[0] = e.dim[1] = e.dim[2] = 1;
    e.start = ws_off[f * P.C]; e.n = ws_off[(f + 1) * P.C] - e.start; e.n0 = e.n;
    e.thr = 0.0; e.passthrough = 0; e.cols = 0; e.kind = e.n > 0 ? kGrpPair : kGrpEmpty; e.cls = f; e.skip = 0; e.seg = -1;
    geom[f] = e;
}
__global__ void __launch_bounds__(kFbThreads) fb_ws_rows_kernel(const FbParams P, const float *__restrict__ ws_xyz, const int *__restrict__ ws_off,
                                                                const int *n_dev, double *__restrict__ pd, int *__restrict__ grp) {
    const int n = *n_dev;
    const long long stride = (long long)gridDim.x * kFbThreads;
    for (long long i = (long long)blockIdx.x * kFbThreads + threadIdx.x; i < n; i += stride) {
        pd[3 * i] = (double)ws_xyz[3 * i]; pd[3 * i + 1] = (double)ws_xyz[3 * i + 1]; pd[3 * i + 2] = (double)ws_xyz[3 * i + 2];
        grp[i] = fb_frame_of_ws_row(ws_off, P.C, P.F, (int)i);
    }
}
struct FbWsSelectOp {
    using Payload = int;
    const double *avg;
    const int *grp;
    const PairGeom *geom;
    const float *ws_xyz;
    float *filt;
    int *filt_grp;
    __device__ __forceinline__ unsigned eval4(long long i0, long long n, int (&p)[4]) const {
        unsigned keep = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const long long i = i0 + j;
            if (i >= n) break;
            const double a = avg[i];
            if (a > 0.0 && a < geom[grp[i]].thr) { p[j] = (int)i; keep |= 1u << j; }
        }
        return keep;
    }
    __device__ __forceinline__ void store(long long pos, const int &i) const {
        filt[3 * pos] = ws_xyz[3 * (long long)i]; filt[3 * pos + 1] = ws_xyz[3 * (long long)i + 1]; filt[3 * pos + 2] = ws_xyz[3 * (long long)i + 2];
        filt_grp[pos] = grp[i];
    }
};

// --------------------------------------------------------------------------------- host side
struct FbPlan {
    FbCounts *cnt;
    void *sort_state[3];
    unsigned long long *sel_state[4];
    size_t zero_off, zero_bytes;   // region re-zeroed at every run
    unsigned *bm0, *bm1;           // persistent bitmaps (zeroed once by init, self-cleaning afterwards)
    CullCell *cull;                // persistent depth-interval cull table (built once by init)
    unsigned *cull_meta;           // [0] = largest interval end of the table (float bits), read back by rt3d_frame_batch_init
    unsigned long long *keys[2];
    unsigned *vals[2];
    float *obj_pts, *fused32;
    float4 *obj_sorted;
    unsigned char *keep;
    int *frame_cnt;
    SubGrid *sgrid;
    SubCell *fine, *super, *leaf;
    SubRep *fine_rep, *super_rep;
    unsigned long long *obj_key;
    int *seg_off, *fused_grp, *slab_cnt, *img_total;
    int *seg_dst, *seg_group, *grp;   // grouping: where a segment's rows go in group order, its group, the group of every row
    unsigned long long *gcode;        // voxel indices of every row, group order
    LatGeom *lat;                     // bit planes of the lattice neighbour search (lat_words == 0: not used)
    unsigned *lat_bits, *lat_cursor;
    unsigned char *lat_need;
    size_t lat_words;
    double *pd, *ps, *avg;
    PairGeom *geom;
    int *col_tab;
    // RT3D_FUSE_CENTROID only
    double *filt64, *cen;             // filtered rows of every group (group order), group centroids
    int *filt_off, *slot_n;
    FbOutCloud *out;
    unsigned *masks_u8;               // RT3D_F32 masks converted to bytes
    // ws_sor only: a second set of outlier-filter buffers sized for the workspace
    double *ws_pd, *ws_ps, *ws_avg;
    int *ws_grp, *ws_filt_grp, *ws_filt_off, *ws_coltab;
    unsigned long long *ws_keys[2];
    unsigned *ws_vals[2];
    void *ws_sort_state;
    unsigned long long *ws_sel_state;
    PairGeom *ws_geom;
    float *ws_filt;
    size_t total;
    bool ok;
};

static int ceil_log2(long long x) { int b = 0; while ((1ll << b) < x) ++b; return b; }

enum { kPlanInit = 1, kPlanFastDiv = 2, kPlanCullCold = 4 };   // rt3d_frame_config::plan bits

struct FbHost {
    FbParams P;
    int seg_bits, group_bits, frame_bits, obj_total_bits, sor_total_bits, sub_total_bits;
    long long ws_cap, obj_cap, fused_cap;
    int nseg, ngroups;
    bool mask_f32, ws_sor;
    int ws_sor_bits;
    size_t lat_words;   // budget of the bit planes in 32-bit words; 0: lattice neighbour search not applicable
};

static int fb_setup(const rt3d_frame_config *cfg, FbHost &h) {
    RT3D_REQUIRE(cfg != nullptr, RT3D_EINVAL, "frame batch: null config");
    RT3D_REQUIRE(cfg->F >= 1 && cfg->C >= 1 && cfg->C <= kFbMaxCams && cfg->H >= 1 && cfg->W >= 4 && cfg->n_obj >= 0,
                 RT3D_EINVAL, "frame batch: bad F/C/H/W/n_obj (C <= %d)", kFbMaxCams);
    RT3D_REQUIRE(cfg->W % 4 == 0, RT3D_EINVAL, "frame batch: image width must be a multiple of 4 (got %d)", cfg->W);
    RT3D_REQUIRE(cfg->voxel > 0 && cfg->max_range > 0, RT3D_EINVAL, "frame batch: voxel and max_range must be positive");
    RT3D_REQUIRE(cfg->n_obj == 0 || (cfg->sor_k >= 1 && cfg->sor_k <= 64 && cfg->sor_std_ratio > 0), RT3D_EINVAL,
                 "[RemoveStatisticalOutliers] Illegal input parameters, the number of neighbors and standard "
                 "deviation ratio must be positive (and nb_neighbors <= 64).");
    RT3D_REQUIRE(cfg->fuse_mode == RT3D_FUSE_BY_INDEX || cfg->fuse_mode == RT3D_FUSE_CENTROID, RT3D_EINVAL, "frame batch: unknown fuse_mode %d", cfg->fuse_mode);
    RT3D_REQUIRE(cfg->fuse_mode == RT3D_FUSE_BY_INDEX || (cfg->C == 2 && cfg->n_obj <= 32), RT3D_EINVAL,
                 "frame batch: RT3D_FUSE_CENTROID is fuse_point_clouds_centroid: exactly 2 cameras, at most 32 masks per camera");
    RT3D_REQUIRE(cfg->mask_dtype == RT3D_U8 || cfg->mask_dtype == RT3D_F32, RT3D_EINVAL, "frame batch: mask_dtype must be RT3D_U8 or RT3D_F32");
    FbParams &P = h.P;
    P.F = cfg->F; P.C = cfg->C; P.H = cfg->H; P.W = cfg->W; P.n_obj = cfg->n_obj;
    P.fuse_mode = cfg->fuse_mode; P.fuse_thr = cfg->fuse_thr;
    P.slots = cfg->fuse_mode == RT3D_FUSE_CENTROID ? 2 * cfg->n_obj : cfg->n_obj;
    h.mask_f32 = cfg->mask_dtype == RT3D_F32 && cfg->n_obj > 0;
    h.ws_sor = cfg->ws_sor != 0;
    RT3D_REQUIRE(!h.ws_sor || (cfg->sor_k >= 1 && cfg->sor_k <= 64 && cfg->sor_std_ratio > 0), RT3D_EINVAL,
                 "[RemoveStatisticalOutliers] Illegal input parameters, the number of neighbors and standard "
                 "deviation ratio must be positive (and nb_neighbors <= 64).");
    // the IEEE flavour may use the 3-instruction division once rt3d_frame_batch_init has verified it for each
    // divisor (exhaustively, on the device; recorded in cfg->plan) and the camera parameters keep every dividend
    // in the verified range
    bool bounded = true;
    for (int c = 0; c < cfg->C; ++c) {
        for (int i = 0; i < 9; ++i) bounded = bounded && fabsf(cfg->Rt[c][i]) <= 1048576.0f;
        for (int i = 0; i < 3; ++i) bounded = bounded && fabsf(cfg->Rt[c][9 + i]) <= 1.0e12f;
        for (int i = 0; i < 2; ++i) {   // principal point: zero or in [2^-20, 2^20] (see depth_in_fast_range)
            const double a = fabs((double)(float)cfg->intr[c][i]);
            bounded = bounded && (a == 0.0 || (a >= 0x1p-20 && a <= 1048576.0));
        }
    }
    bounded = bounded && cfg->W <= 8192 && cfg->H <= 8192 && cfg->voxel >= 0x1p-30;
    for (int i = 0; i < 6; ++i) bounded = bounded && fabsf(cfg->box[i]) <= 1048576.0f;
    for (int i = 0; i < 6; ++i) bounded = bounded && fabs((double)cfg->box[i]) / cfg->voxel < 2097000.0;   // in-box voxel indices below 2^21 (fb_pixel_pair's rounding)
    const bool all_fast = (cfg->flags & 1) == RT3D_DIV_IEEE && bounded && (cfg->plan & kPlanFastDiv) != 0;
    auto make_div = [&](double b) { return all_fast ? DivBy::fast_ieee(b) : DivBy(b, cfg->flags); };
    for (int c = 0; c < cfg->C; ++c) {
        P.cam[c].cx = (float)cfg->intr[c][0]; P.cam[c].cy = (float)cfg->intr[c][1];
        P.cam[c].dfx = make_div(cfg->intr[c][2]); P.cam[c].dfy = make_div(cfg->intr[c][3]);
        for (int i = 0; i < 9; ++i) P.cam[c].rt.R[i] = cfg->Rt[c][i];
        for (int i = 0; i < 3; ++i) P.cam[c].rt.t[i] = cfg->Rt[c][9 + i];
    }
    for (int a = 0; a < 3; ++a) { P.box.lo[a] = cfg->box[2 * a]; P.box.hi[a] = cfg->box[2 * a + 1]; }
    P.div_v = make_div(cfg->voxel);
    P.v32 = (float)cfg->voxel;
    long long nvox = 1;
    for (int a = 0; a < 3; ++a) {
        // voxel indices are monotone in the coordinate, so the box corners bound them
        volatile float qlo = P.div_v.mode == 1 ? P.box.lo[a] * P.div_v.inv_b : P.box.lo[a] / P.div_v.b;
        volatile float qhi = P.div_v.mode == 1 ? P.box.hi[a] * P.div_v.inv_b : P.box.hi[a] / P.div_v.b;
        const long long klo = (long long)nearbyintf(qlo), khi = (long long)nearbyintf(qhi);
        RT3D_REQUIRE(khi >= klo && khi - klo < (1ll << 20), RT3D_EINVAL, "frame batch: crop box / voxel size out of range");
        P.k0[a] = (int)klo; P.nk[a] = (int)(khi - klo + 1);
        nvox *= P.nk[a];
    }
    RT3D_REQUIRE(nvox < (1ll << 31), RT3D_ERANGE, "frame batch: %lld voxels in the crop box exceed 2^31 (use a larger voxel)", nvox);
    // occupancy bitmap: one 32-bit word per (x, y group of 32, z); level 1 = one bit per word
    P.nyg = (P.nk[1] + 31) / 32;
    P.nzp = (P.nk[2] + 31) / 32 * 32;
    RT3D_REQUIRE(P.nzp <= 65535, RT3D_ERANGE, "frame batch: crop box too tall for this voxel size");
    P.slabs_per_img = P.nk[0] * P.nyg;
    P.tiles_per_img = (P.slabs_per_img + 7) / 8;
    const long long words0 = (long long)P.slabs_per_img * P.nzp;
    RT3D_REQUIRE(words0 < (1ll << 31), RT3D_ERANGE, "frame batch: occupancy bitmap too large (use a larger voxel)");
    P.words0_per_img = (int)words0; P.words1_per_img = (int)(words0 / 32);   // level 0: z-major planes of slabs_per_img words; level 1: slab-major
    RT3D_REQUIRE((long long)P.tiles_per_img * cfg->F * cfg->C < (1ll << 30), RT3D_ERANGE, "frame batch: too many bitmap tiles");
    const long long kmax = (long long)ceil(cfg->max_range / cfg->voxel) + 2;
    P.obits = ceil_log2(2 * kmax); P.obias = 1 << (P.obits - 1);
    RT3D_REQUIRE(P.obits <= 20, RT3D_EINVAL, "frame batch: max_range / voxel too large");
    P.seg_shift = 3 * P.obits;
    h.nseg = cfg->F * cfg->n_obj * cfg->C; h.ngroups = cfg->F * P.slots;
    h.seg_bits = ceil_log2(h.nseg > 1 ? h.nseg : 2); h.group_bits = ceil_log2(h.ngroups > 1 ? h.ngroups : 2);
    h.frame_bits = ceil_log2(cfg->F > 1 ? cfg->F : 2);
    h.obj_total_bits = P.seg_shift + h.seg_bits;
    h.sor_total_bits = kFbCodeBits + h.group_bits;
    h.sub_total_bits = subtract_key_bits(cfg->F);
    h.ws_sor_bits = kFbCodeBits + h.frame_bits;
    RT3D_REQUIRE(h.obj_total_bits <= 64, RT3D_ERANGE, "frame batch: object keys need %d bits", h.obj_total_bits);
    const long long groups = (long long)cfg->F * cfg->C * ((long long)cfg->H * cfg->W / 4);
    RT3D_REQUIRE(groups < (1ll << 30) && (long long)cfg->H * cfg->W < (1ll << 31), RT3D_ERANGE, "frame batch: too many pixels per batch (F * C * H * W must stay below 2^32)");
    const unsigned pj_cx = (unsigned)((cfg->W + kPjChunk * kPjCols - 1) / (kPjChunk * kPjCols)), pj_ty = (unsigned)((cfg->H + kPjRows - 1) / kPjRows);
    RT3D_REQUIRE((long long)pj_cx * pj_ty * cfg->F * cfg->C < (1ll << 31), RT3D_ERANGE, "frame batch: too many image tiles per batch");
    RT3D_REQUIRE((long long)cfg->C * (pj_cx * kPjChunk * (kPjCols / kCullCell)) * ((pj_ty + 1) * (kPjRows / kCullCell)) < (1ll << 31), RT3D_ERANGE,
                 "frame batch: image too large for the cull table");
    P.pj_chunks = pj_cx * pj_ty * (unsigned)(cfg->F * cfg->C);
    const unsigned tm_ty = (unsigned)((cfg->H + kTmRows - 1) / kTmRows);
    P.tm_chunks = pj_cx * tm_ty * (unsigned)(cfg->F * cfg->C);
    P.div_tm_cpi = make_fastdiv(pj_cx * tm_ty);
    P.div_pj_cpi = make_fastdiv(pj_cx * pj_ty);
    P.div_pj_cx = make_fastdiv(pj_cx);
    P.div_c = make_fastdiv((unsigned)cfg->C);
    double cell_factor = 3.0;
#ifdef RT3D_PROFILING   // profiling builds only: knobs from the environment (never in the shipped library)
    P.pj_debug = 0;
    if (const char *e = getenv("RT3D_PJ_DEBUG")) P.pj_debug = atoi(e);
    if (const char *e = getenv("RT3D_SOR_CELL_VOXELS")) { const double v = atof(e); if (v > 0.5 && v < 64.0) cell_factor = v; }
#endif

    P.div_nyg = make_fastdiv((unsigned)P.nyg);
    P.div_spi = make_fastdiv((unsigned)P.slabs_per_img);
    RT3D_REQUIRE((long long)P.slabs_per_img * cfg->F * cfg->C < (1ll << 31), RT3D_ERANGE, "frame batch: too many bitmap slabs per batch");
    P.cull_cx = (int)pj_cx * kPjChunk * (kPjCols / kCullCell); P.cull_cy = (int)tm_ty * (kTmRows / kCullCell);   // whole stages of the tensor-map kernel (>= whole 8-row tiles)
    P.word_x = P.nyg * P.nzp;
    P.force_cold = (cfg->plan & kPlanCullCold) != 0;
    // level-0 words are stored plane by plane (z major): the occupied voxels of a mostly horizontal scene then
    // sit in a few contiguous planes instead of one 32-byte sector every nzp words
    P.word0_z = P.slabs_per_img;
#if RT3D_BM_XFAST
    P.word0_x = 1; P.word0_y = P.nk[0];
#else
    P.word0_x = P.nyg; P.word0_y = 1;
#endif
    P.word0_c = (int)(0u - (unsigned)P.k0[2] * (unsigned)P.word0_z - (unsigned)P.k0[0] * (unsigned)P.word0_x);
    P.word_c = (int)(0u - (unsigned)P.k0[0] * (unsigned)P.word_x - (unsigned)P.k0[2]);   // modulo 2^32, like the device side
    P.word_cm = (int)((unsigned)P.word_c - kRoundMagicBits * (unsigned)P.word_x - kRoundMagicBits);
    P.word0_cm = (int)((unsigned)P.word0_c - kRoundMagicBits * (unsigned)P.word0_z - kRoundMagicBits * (unsigned)P.word0_x);
    P.k0y_m = (int)((unsigned)P.k0[1] + kRoundMagicBits);
    P.sor_cell = cell_factor * cfg->voxel * sqrt((cfg->sor_k > 0 ? cfg->sor_k : 20) / 20.0);
    P.sor_k = cfg->sor_k; P.sor_ratio = cfg->sor_std_ratio;
    P.sub_t2 = (double)cfg->sub_thr * (double)cfg->sub_thr; P.sub_t2f = (float)P.sub_t2;
    h.ws_cap = cfg->ws_capacity; h.obj_cap = cfg->obj_capacity;
    h.fused_cap = cfg->fused_capacity > 0 ? cfg->fused_capacity : cfg->obj_capacity;
    RT3D_REQUIRE(h.ws_cap >= 1 && h.ws_cap < (1ll << 30) && h.obj_cap >= 1 && h.obj_cap < (1ll << 30) && h.fused_cap < (1ll << 30), RT3D_ERANGE,
                 "frame batch: capacities must be in [1, 2^30)");
    RT3D_REQUIRE(cfg->fuse_mode == RT3D_FUSE_CENTROID || h.fused_cap >= h.obj_cap, RT3D_EINVAL,
                 "frame batch: fused_capacity below obj_capacity (RT3D_FUSE_BY_INDEX writes up to obj_capacity rows)");
    // bit-plane neighbour search in front of the grid search (RT3D_SOR_GRID_ONLY switches it off; same results either way):
    // at most two clouds per group (one duplicate per voxel), k <= 32 candidates per query;
    // 64 bytes of planes per object point capacity (a surface cloud needs ~5-12 at the 3-sigma core box), capped at 2^31 bits
    h.lat_words = 0;
    // automatic: on for large batches (four more launches cost ~60 us; the planes save ~0.28 us per thousand queries) of
    // voxels no finer than 3 mm (measured at 2 mm, HD720: half the queries settle and the 350 MB of planes cost more than that saves)
    const bool lat_auto = h.obj_cap >= 1000000 && cfg->voxel >= 0.003, lat_on = (cfg->flags & RT3D_SOR_BITPLANES) || (lat_auto && !(cfg->flags & RT3D_SOR_GRID_ONLY));
    if (lat_on && cfg->n_obj > 0 && cfg->sor_k <= 32 && (cfg->fuse_mode == RT3D_FUSE_CENTROID || cfg->C <= 2)) {
        // 256 bytes of planes per object-point capacity at most (5 mm surface clouds use 5-12, 2 mm ones ~100); only the
        // words the groups actually claim are zeroed per run
        unsigned long long w = 64ull * (unsigned long long)h.obj_cap;
        if (w > (1ull << 28)) w = 1ull << 28;
        if (w < 4096ull) w = 4096ull;
        h.lat_words = (size_t)w;
    }
    return RT3D_OK;
}

static FbPlan fb_plan(const FbHost &h, void *scratch, size_t bytes) {
    Arena a(scratch, bytes);
    FbPlan p;
    const FbParams &P = h.P;
    const size_t nimg = (size_t)P.F * P.C;
    // persistent region first: the bitmaps
    p.bm0 = a.take<unsigned>(nimg * P.words0_per_img);
    p.bm1 = a.take<unsigned>(nimg * P.words1_per_img);
    p.cull = a.take<CullCell>((size_t)P.C * (size_t)P.cull_cx * (size_t)P.cull_cy);
    p.cull_meta = a.take<unsigned>(4);
    // per-run zeroed region
    p.zero_off = a.off;
    p.cnt = a.take<FbCounts>(1);
    p.sort_state[0] = a.take<char>(sort_state_bytes(h.obj_cap, h.obj_total_bits));
    p.sort_state[1] = a.take<char>(sort_state_bytes(h.obj_cap, h.sor_total_bits));
    p.sort_state[2] = a.take<char>(sort_state_bytes(h.fused_cap, h.sub_total_bits));
    p.sel_state[0] = a.take<unsigned long long>(1);
    p.sel_state[1] = a.take<unsigned long long>(select_state_words(h.obj_cap));
    p.sel_state[2] = a.take<unsigned long long>(select_state_words(h.obj_cap));
    p.sel_state[3] = a.take<unsigned long long>(select_state_words(h.ws_cap));
    p.sgrid = a.take<SubGrid>((size_t)P.F + 1);    // zeroed: a frame without objects has grid.n == 0
    p.lat_cursor = a.take<unsigned>(4);            // zeroed: words of bit planes claimed so far
    p.ws_sort_state = nullptr; p.ws_sel_state = nullptr;
    if (h.ws_sor) {
        p.ws_sort_state = a.take<char>(sort_state_bytes(h.ws_cap, h.ws_sor_bits));
        p.ws_sel_state = a.take<unsigned long long>(select_state_words(h.ws_cap));
    }
    p.zero_bytes = a.off - p.zero_off;
    p.slab_cnt = a.take<int>(nimg * (size_t)P.slabs_per_img + 1);
    p.img_total = a.take<int>(nimg + 1);
    const size_t kcap = (size_t)(h.obj_cap > h.fused_cap ? h.obj_cap : h.fused_cap) + 4;   // the subtraction sorts the fused rows
    p.keys[0] = a.take<unsigned long long>(kcap);
    p.keys[1] = a.take<unsigned long long>(kcap);
    p.vals[0] = a.take<unsigned>(kcap);
    p.vals[1] = a.take<unsigned>(kcap);
    p.obj_pts = a.take<float>(3 * (size_t)h.obj_cap + 4);
    p.fused32 = a.take<float>(3 * (size_t)h.fused_cap + 4);
    p.obj_sorted = a.take<float4>((size_t)h.fused_cap + 1);
    p.obj_key = a.take<unsigned long long>((size_t)h.obj_cap + 4);
    p.seg_off = a.take<int>((size_t)h.nseg + 2);
    p.seg_dst = a.take<int>((size_t)h.nseg + 2);
    p.seg_group = a.take<int>((size_t)h.nseg + 2);
    p.grp = a.take<int>((size_t)h.obj_cap + 4);
    p.gcode = a.take<unsigned long long>((size_t)h.obj_cap + 4);
    p.lat_words = h.lat_words;
    p.lat = a.take<LatGeom>((size_t)h.ngroups + 1);
    p.lat_need = a.take<unsigned char>((size_t)h.obj_cap + 16);
    p.lat_bits = a.take<unsigned>(h.lat_words + 4);
    p.fused_grp = a.take<int>((size_t)h.obj_cap + 4);
    p.keep = a.take<unsigned char>((size_t)h.ws_cap + 4);
    p.frame_cnt = a.take<int>((size_t)P.F + 1);
    p.pd = a.take<double>(3 * (size_t)h.obj_cap + 4);
    p.ps = a.take<double>(3 * (size_t)h.obj_cap + 4);
    p.avg = a.take<double>((size_t)h.obj_cap + 4);
    p.geom = a.take<PairGeom>((size_t)h.ngroups + 1);
    p.col_tab = a.take<int>((size_t)h.ngroups * (kFbColCap + 1));
    p.filt64 = nullptr; p.cen = nullptr; p.filt_off = nullptr; p.slot_n = nullptr; p.out = nullptr;
    if (P.fuse_mode == RT3D_FUSE_CENTROID) {
        p.filt64 = a.take<double>(3 * (size_t)h.obj_cap + 4);
        p.cen = a.take<double>(3 * (size_t)h.ngroups + 4);
        p.filt_off = a.take<int>((size_t)h.ngroups + 2);
        p.slot_n = a.take<int>((size_t)h.ngroups + 2);
        p.out = a.take<FbOutCloud>((size_t)h.ngroups + 1);
    }
    p.masks_u8 = h.mask_f32 ? a.take<unsigned>(nimg * (size_t)P.n_obj * ((size_t)P.H * P.W / 4) + 4) : nullptr;
    p.ws_pd = p.ws_ps = p.ws_avg = nullptr; p.ws_grp = p.ws_filt_grp = p.ws_filt_off = p.ws_coltab = nullptr;
    p.ws_keys[0] = p.ws_keys[1] = nullptr; p.ws_vals[0] = p.ws_vals[1] = nullptr; p.ws_geom = nullptr; p.ws_filt = nullptr;
    if (h.ws_sor) {
        const size_t wc = (size_t)h.ws_cap + 4;
        p.ws_pd = a.take<double>(3 * wc); p.ws_ps = a.take<double>(3 * wc); p.ws_avg = a.take<double>(wc);
        p.ws_grp = a.take<int>(wc); p.ws_filt_grp = a.take<int>(wc); p.ws_filt_off = a.take<int>((size_t)P.F + 2);
        p.ws_coltab = a.take<int>((size_t)P.F * (kFbColCap + 1));
        p.ws_keys[0] = a.take<unsigned long long>(wc); p.ws_keys[1] = a.take<unsigned long long>(wc);
        p.ws_vals[0] = a.take<unsigned>(wc); p.ws_vals[1] = a.take<unsigned>(wc);
        p.ws_geom = a.take<PairGeom>((size_t)P.F + 1);
        p.ws_filt = a.take<float>(3 * wc);
    }
    p.fine = a.take<SubCell>((size_t)P.F * kSubFineCells);
    p.leaf = a.take<SubCell>((size_t)P.F * kSubLeafCells);
    p.super = a.take<SubCell>((size_t)P.F * kSubSuperCells);
    p.fine_rep = a.take<SubRep>((size_t)P.F * kSubFineCells);
    p.super_rep = a.take<SubRep>((size_t)P.F * kSubSuperCells);
    p.total = a.off;
    p.ok = a.ok();
    return p;
}

// optional per-stage CUDA events (bench.py: per-stage p50 ms and the roofline of the project kernel)
struct StageTimer {
    cudaEvent_t ev[6];   // boundaries: start, after A, after B, after C, after D, after E
    bool recorded;
};

// ---- tensor maps of the TMA project kernel.  cuTensorMapEncodeTiled is a driver-API entry point: fetched once
// through the runtime (no link-time dependency on libcuda); encoding is host arithmetic on the pointers of this call.
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);
static EncodeTiledFn tensor_map_encoder() {
    static EncodeTiledFn fn = []() -> EncodeTiledFn {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
        return (EncodeTiledFn)p;
    }();
    return fn;
}

// depth  [nimg][H][W] fp32 seen as {16 px, H rows, W/16 patches, nimg}: a box {16, rows, 4, 1} lands patch-major, i.e. as
//        [patch][row][16 px] -- lane (r, cg) of a warp then reads its float4 of patch q at q * 512 + lane * 16: every
//        quarter-warp covers 128 contiguous bytes (no bank conflict), and the stage layout equals the cp.async kernel's.
// masks  [nimg * n_obj][H][W] bytes, the same view with 1-byte elements (16-byte inner box).
// cull   [C][cull_cy][cull_cx] cells of 8 bytes seen as 32-bit words: box {32, rows / 4, 1} = the rows of 16 cells under a stage.
static int fb_encode_maps(const FbParams &P, const float *depth, const unsigned char *masks, const CullCell *cull, TmMaps &M) {
    EncodeTiledFn enc = tensor_map_encoder();
    RT3D_REQUIRE(enc != nullptr, RT3D_ECUDA, "frame batch: cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t W = (cuuint64_t)P.W, H = (cuuint64_t)P.H, nimg = (cuuint64_t)P.F * P.C;
    const cuuint32_t ones[4] = {1, 1, 1, 1};
    {
        const cuuint64_t dim[4] = {16, H, W / 16, nimg}, str[3] = {W * 4, 64, H * W * 4};
        const cuuint32_t box[4] = {16, kTmRows, kPjCols / 16, 1};
        const CUresult rc = enc(&M.depth, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void *)depth, dim, str, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        RT3D_REQUIRE(rc == CUDA_SUCCESS, RT3D_ECUDA, "frame batch: encoding the depth tensor map failed (CUresult %d)", (int)rc);
    }
    if (P.n_obj > 0) {
        const cuuint64_t dim[4] = {16, H, W / 16, nimg * (cuuint64_t)P.n_obj}, str[3] = {W, 16, H * W};
        const cuuint32_t box[4] = {16, kTmRows, kPjCols / 16, 1};
        const CUresult rc = enc(&M.masks, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, (void *)masks, dim, str, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        RT3D_REQUIRE(rc == CUDA_SUCCESS, RT3D_ECUDA, "frame batch: encoding the mask tensor map failed (CUresult %d)", (int)rc);
    } else {
        M.masks = M.depth;
    }
    {
        const cuuint64_t dim[3] = {2 * (cuuint64_t)P.cull_cx, (cuuint64_t)P.cull_cy, (cuuint64_t)P.C};
        const cuuint64_t str[2] = {8 * (cuuint64_t)P.cull_cx, 8 * (cuuint64_t)P.cull_cx * (cuuint64_t)P.cull_cy};
        const cuuint32_t box[3] = {32, kTmRows / kCullCell, 1};
        const CUresult rc = enc(&M.cull, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, (void *)cull, dim, str, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        RT3D_REQUIRE(rc == CUDA_SUCCESS, RT3D_ECUDA, "frame batch: encoding the cull-table tensor map failed (CUresult %d)", (int)rc);
    }
    return RT3D_OK;
}

static int grid_for_n(long long n, int threads, int per_sm);

// The grid search of the outlier filter for one batch of groups (objects: stage D; workspace: stage B'): grid geometry,
// cell keys, sort, column tables, k-NN mean distances, sequential mean / sigma -> geom[g].thr and avg[] are final.
struct SorCore {
    const double *pd; const int *grp; const int *n_dev; long long cap;
    PairGeom *geom; int ngroups, total_bits;
    unsigned long long *keys[2]; unsigned *vals[2]; void *sort_state;
    double *ps, *avg; int *col_tab, *ticket, *status;
    const unsigned long long *skeys; const unsigned *svals;   // out: the sorted keys / payload
};
static int launch_sor_core(const FbParams &P, SorCore &S, const unsigned char *need, cudaStream_t stream);

static int grid_for_n(long long n, int threads, int per_sm) { return persistent_blocks((n + threads - 1) / threads, per_sm); }

static int launch_sor_core(const FbParams &P, SorCore &S, const unsigned char *need, cudaStream_t stream) {
    const int g = grid_for_n(S.cap, kFbThreads, 8);
    fb_sor_prepare_kernel<<<S.ngroups, 128, 0, stream>>>(P, S.pd, S.geom);
    const int passes = sort_max_passes(S.total_bits);
    fb_sor_cellkey_kernel<<<g, kFbThreads, 0, stream>>>(P, S.pd, S.grp, S.n_dev, S.geom, passes, S.keys[0], S.vals[0], sort_hist_ptr(S.sort_state));
    RT3D_LAUNCH_OK();
    SortArgs sa;
    sa.keys[0] = S.keys[0]; sa.keys[1] = S.keys[1]; sa.vals[0] = S.vals[0]; sa.vals[1] = S.vals[1];
    sa.n = S.cap; sa.n_dev = S.n_dev; sa.total_bits = S.total_bits; sa.total_bits_dev = nullptr;
    sa.state = S.sort_state; sa.status = S.status; sa.hist_done = true;
    const int rc = launch_sort(sa, stream);
    if (rc) return rc;
    S.skeys = S.keys[passes & 1]; S.svals = S.vals[passes & 1];
    fb_gather_f64_kernel<<<g, kFbThreads, 0, stream>>>(S.pd, S.n_dev, S.svals, S.ps);
    fb_sor_cols_kernel<<<g, kFbThreads, 0, stream>>>(S.skeys, S.n_dev, S.geom, S.col_tab);
    RT3D_LAUNCH_OK();
    return RT3D_OK;
}
static int launch_sor_knn(const FbParams &P, const SorCore &S, const unsigned char *need, cudaStream_t stream) {
    const int kblocks = persistent_blocks((S.cap * 32 + kFbThreads - 1) / kFbThreads, 8);
    // (the mask test is a template parameter: as a run-time test it changed the register allocation of the 40-register
    // kernel enough to spill in its hot loop -- 2.31 vs 1.72 ms with the mask absent; the masked variant gets 64 registers)
    if (need != nullptr) fb_sor_knn_kernel<1, true><<<kblocks, kFbThreads, 0, stream>>>(S.ps, S.skeys, S.svals, S.n_dev, P.sor_k, S.geom, S.col_tab, S.avg, S.ticket, need);
    else if (P.sor_k <= 32) fb_sor_knn_kernel<1, false><<<kblocks, kFbThreads, 0, stream>>>(S.ps, S.skeys, S.svals, S.n_dev, P.sor_k, S.geom, S.col_tab, S.avg, S.ticket, nullptr);
    else fb_sor_knn_kernel<2, false><<<kblocks, kFbThreads, 0, stream>>>(S.ps, S.skeys, S.svals, S.n_dev, P.sor_k, S.geom, S.col_tab, S.avg, S.ticket, nullptr);
    fb_sor_stats_kernel<<<S.ngroups, kFbThreads, 0, stream>>>(S.avg, P.sor_ratio, S.geom);
    RT3D_LAUNCH_OK();
    return RT3D_OK;
}

}  // namespace rt3d

using namespace rt3d;

extern "C" int rt3d_frame_batch_slots(const rt3d_frame_config *cfg) {
    if (cfg == nullptr) return 0;
    return cfg->fuse_mode == RT3D_FUSE_CENTROID ? 2 * cfg->n_obj : cfg->n_obj;
}

extern "C" size_t rt3d_frame_batch_scratch_bytes(const rt3d_frame_config *cfg) {
    FbHost h;
    if (fb_setup(cfg, h) != RT3D_OK) return 0;
    return fb_plan(h, nullptr, 0).total;
}

extern "C" int rt3d_frame_batch_init(rt3d_frame_config *cfg, void *scratch, size_t scratch_bytes, rt3d_stream_t stream_) {
    RT3D_REQUIRE(cfg != nullptr, RT3D_EINVAL, "frame batch: null config");
    cfg->plan = 0;
    FbHost h;
    int rc = fb_setup(cfg, h);
    if (rc) return rc;
    RT3D_REQUIRE(scratch != nullptr, RT3D_EINVAL, "frame batch: null scratch");
    FbPlan p = fb_plan(h, scratch, scratch_bytes);
    RT3D_REQUIRE(p.ok, RT3D_ESCRATCH, "scratch too small: need %zu bytes, have %zu", p.total, scratch_bytes);
    // the occupancy bitmaps must start out empty; every run leaves them empty again
    RT3D_CUDA_OK(cudaMemsetAsync(scratch, 0, p.zero_off, (cudaStream_t)stream_));
    {
        const long long cells = (long long)h.P.C * h.P.cull_cx * h.P.cull_cy;
        fb_cull_build_kernel<<<grid_for_n(cells, 256, 8), 256, 0, (cudaStream_t)stream_>>>(h.P, p.cull, p.cull_meta);
        RT3D_LAUNCH_OK();
    }
    {   // the lattice offsets in constant memory (idempotent; per device)
        unsigned short host_off[128];
        RT3D_CUDA_OK(cudaMemcpyFromSymbol(host_off, kLatOff, sizeof host_off));
        RT3D_CUDA_OK(cudaMemcpyToSymbolAsync(c_lat_off, host_off, sizeof host_off, 0, cudaMemcpyHostToDevice, (cudaStream_t)stream_));
    }
    // IEEE flavour: qualify the fast division for this configuration's divisors (exhaustive device check, cached
    // per divisor; if any divisor fails, the configuration keeps the plain IEEE division).  The outcome travels in
    // cfg->plan, so run() neither consults the cache nor depends on what other configurations did.
    int plan = kPlanInit;
    if ((cfg->flags & 1) == RT3D_DIV_IEEE) {
        bool ok = verify_fast_division(cfg->voxel, (cudaStream_t)stream_);
        for (int c = 0; c < cfg->C; ++c)
            ok = verify_fast_division(cfg->intr[c][2], (cudaStream_t)stream_) && verify_fast_division(cfg->intr[c][3], (cudaStream_t)stream_) && ok;
        if (ok) plan |= kPlanFastDiv;
    }
    RT3D_CUDA_OK(cudaStreamSynchronize((cudaStream_t)stream_));
    {   // depth range of the cull table: outside the verified range every candidate takes the plain IEEE path
        // (beyond 2^70 m every candidate of the configuration takes the out-of-line IEEE division.  The lower end stays a
        // per-pixel test: the cameras of the reference rig stand inside its crop box, so depths down to zero are in-box.)
        unsigned hb = 0u;
        RT3D_CUDA_OK(cudaMemcpy(&hb, p.cull_meta, sizeof hb, cudaMemcpyDeviceToHost));
        if (hb > kFastHiBits) plan |= kPlanCullCold;
    }
    cfg->plan = plan;
    return RT3D_OK;
}

extern "C" rt3d_stage_timer *rt3d_stage_timer_create(void) {
    StageTimer *t = new StageTimer();
    t->recorded = false;
    for (int i = 0; i < 6; ++i) {
        if (cudaEventCreate(&t->ev[i]) != cudaSuccess) { set_error("cudaEventCreate failed"); delete t; return nullptr; }
    }
    return (rt3d_stage_timer *)t;
}

extern "C" void rt3d_stage_timer_destroy(rt3d_stage_timer *timer) {
    StageTimer *t = (StageTimer *)timer;
    if (!t) return;
    for (int i = 0; i < 6; ++i) cudaEventDestroy(t->ev[i]);
    delete t;
}

extern "C" int rt3d_stage_timer_read(rt3d_stage_timer *timer, float *ms5) {
    StageTimer *t = (StageTimer *)timer;
    RT3D_REQUIRE(t && ms5 && t->recorded, RT3D_EINVAL, "stage timer: nothing recorded");
    RT3D_CUDA_OK(cudaEventSynchronize(t->ev[5]));
    for (int i = 0; i < 5; ++i) RT3D_CUDA_OK(cudaEventElapsedTime(&ms5[i], t->ev[i], t->ev[i + 1]));
    return RT3D_OK;
}

extern "C" int rt3d_frame_batch_run(const rt3d_frame_config *cfg, const float *depth, const void *masks, const int32_t *class_rank,
                                    float *ws_xyz, int32_t *ws_off, double *obj_xyz, int32_t *obj_off, int32_t *obj_meta,
                                    float *kept_xyz, int32_t *kept_off, int32_t *counts, void *scratch, size_t scratch_bytes,
                                    rt3d_stream_t stream_) {
    return rt3d_frame_batch_run_timed(cfg, depth, masks, class_rank, ws_xyz, ws_off, obj_xyz, obj_off, obj_meta, kept_xyz, kept_off,
                                      counts, scratch, scratch_bytes, stream_, nullptr);
}

extern "C" int rt3d_frame_batch_run_timed(const rt3d_frame_config *cfg, const float *depth, const void *masks_, const int32_t *class_rank,
                                          float *ws_xyz, int32_t *ws_off, double *obj_xyz, int32_t *obj_off, int32_t *obj_meta,
                                          float *kept_xyz, int32_t *kept_off, int32_t *counts, void *scratch, size_t scratch_bytes,
                                          rt3d_stream_t stream_, rt3d_stage_timer *timer_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    StageTimer *timer = (StageTimer *)timer_;
#define RT3D_STAGE(i) do { if (timer) RT3D_CUDA_OK(cudaEventRecord(timer->ev[i], stream)); } while (0)
    FbHost h;
    int rc = fb_setup(cfg, h);
    if (rc) return rc;
    const FbParams &P = h.P;
    RT3D_REQUIRE((cfg->plan & kPlanInit) != 0, RT3D_EINVAL, "frame batch: rt3d_frame_batch_init has not been called for this configuration");
    RT3D_REQUIRE(depth && ws_xyz && ws_off && kept_xyz && kept_off && counts && scratch, RT3D_EINVAL, "frame batch: null argument");
    RT3D_REQUIRE(P.n_obj == 0 || (masks_ && obj_xyz && obj_off), RT3D_EINVAL, "frame batch: masks / object outputs required when n_obj > 0");
    RT3D_REQUIRE(P.n_obj == 0 || P.fuse_mode == RT3D_FUSE_BY_INDEX || class_rank, RT3D_EINVAL, "frame batch: RT3D_FUSE_CENTROID needs class_rank");
    RT3D_REQUIRE(((uintptr_t)depth & 15) == 0 && ((uintptr_t)masks_ & 15) == 0, RT3D_EINVAL, "frame batch: depth and masks must be 16-byte aligned");
    FbPlan p = fb_plan(h, scratch, scratch_bytes);
    RT3D_REQUIRE(p.ok, RT3D_ESCRATCH, "scratch too small: need %zu bytes, have %zu", p.total, scratch_bytes);
    RT3D_CUDA_OK(cudaMemsetAsync((char *)scratch + p.zero_off, 0, p.zero_bytes, stream));

    const long long nimg = (long long)P.F * P.C;
    const unsigned char *masks = (const unsigned char *)masks_;
    if (h.mask_f32) {   // fp32 planes -> bytes (one extra pass over the planes; the project kernel reads bytes)
        const long long n4 = nimg * P.n_obj * ((long long)P.H * P.W / 4);
        fb_mask_f32_kernel<<<grid_for_n(n4, kFbThreads, 8), kFbThreads, 0, stream>>>((const float4 *)masks_, n4, p.masks_u8);
        masks = (const unsigned char *)p.masks_u8;
    }
    // ---- A
    RT3D_STAGE(0);
    {
        const int mode = P.div_v.mode == P.cam[0].dfx.mode ? P.div_v.mode : -1;   // all divisors share one flavour
        // TMA tensor copies need 16-byte row strides: image width a multiple of 16 (every ZED resolution); any other
        // width (a multiple of 4) takes the per-lane cp.async kernel
        // The tensor-map pipeline wins on long streams of mostly-background tiles (256 HD720 frame pairs: 0.97 vs 1.10 ms);
        // with few tiles per SM its start-up and drain cost more than it saves (8 HD1080 pairs, 4 masks: 0.27 vs 0.21 ms),
        // and masks beyond the staged ones make the tile time object-bound, where the simpler kernel is ahead (12 masks:
        // 0.63 vs 0.39 ms).  Both kernels run pj_process_tile, so the choice never changes a result.
        const long long tiles = (long long)P.pj_chunks * kPjChunk;
        // (16-row stages: with more than two staged masks a block's ring no longer leaves room for two blocks per SM)
        bool tma = P.W % 16 == 0 && P.n_obj <= 2 && tiles >= 2048ll * kNumSMs;
        if ((cfg->flags & RT3D_PROJECT_MASK) == RT3D_PROJECT_TMA) tma = P.W % 16 == 0;     // explicit choice (tests, tuning)
        if ((cfg->flags & RT3D_PROJECT_MASK) == RT3D_PROJECT_CP_ASYNC) tma = false;
#ifdef RT3D_PROFILING
        if (getenv("RT3D_PJ_NO_TMA") != nullptr) tma = false;
        if (getenv("RT3D_PJ_TMA") != nullptr) tma = P.W % 16 == 0;
#endif
        if (tma) {
            TmMaps maps;
            rc = fb_encode_maps(P, depth, masks, p.cull, maps);
            if (rc) return rc;
            const int pgrid = persistent_blocks(((long long)P.tm_chunks * kPjChunk + kTmConsumers - 1) / kTmConsumers, kTmBlocksPerSM);
            const size_t smem = tm_smem_bytes(P.n_obj);
#define RT3D_PJ(M) do { \
            RT3D_CUDA_OK(cudaFuncSetAttribute(fb_project_tma_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            fb_project_tma_kernel<M><<<pgrid, kTmThreads, smem, stream>>>(P, maps, masks, p.bm0, p.bm1, p.keys[0], (int)h.obj_cap, p.cnt); } while (0)
            if (mode == 2) RT3D_PJ(2); else if (mode == 1) RT3D_PJ(1); else RT3D_PJ(0);
#undef RT3D_PJ
        } else {
            const int pgrid = persistent_blocks((P.pj_chunks + kPjThreads / 32 - 1) / (kPjThreads / 32), kPjBlocksPerSM);
            const size_t pj_smem = kPjCamBytes + (size_t)(kPjThreads / 32) * pj_warp_bytes(P.n_obj);
#define RT3D_PJ(M) do { \
            RT3D_CUDA_OK(cudaFuncSetAttribute(fb_project_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pj_smem)); \
            fb_project_kernel<M><<<pgrid, kPjThreads, pj_smem, stream>>>(P, depth, masks, p.cull, p.bm0, p.bm1, p.keys[0], (int)h.obj_cap, p.cnt); } while (0)
            if (mode == 2) RT3D_PJ(2); else if (mode == 1) RT3D_PJ(1); else RT3D_PJ(0);
#undef RT3D_PJ
        }
        RT3D_LAUNCH_OK();
    }
    RT3D_STAGE(1);   // stage A = the project kernel alone (the roofline kernel)
    fb_clamp_kernel<<<1, 1, 0, stream>>>(p.cnt, (int)h.obj_cap);
    // ---- B
    {
        const int gslab = persistent_blocks((nimg * P.slabs_per_img * 32 + kFbThreads - 1) / kFbThreads, 8);
        fb_ws_count_kernel<<<gslab, kFbThreads, 0, stream>>>(P, p.bm0, p.bm1, p.slab_cnt);
        fb_ws_slab_scan_kernel<<<(int)nimg, kFbThreads, 0, stream>>>(P, p.slab_cnt, p.img_total);
        fb_ws_img_scan_kernel<<<1, kFbThreads, 0, stream>>>(p.img_total, (int)nimg, ws_off, p.cnt);
        fb_ws_emit_kernel<<<gslab, kFbThreads, 0, stream>>>(P, p.bm0, p.bm1, p.slab_cnt, ws_off, ws_xyz, (int)h.ws_cap, p.cnt);
        fb_bm_clear_kernel<<<(int)nimg, kClearThreads, 0, stream>>>(P, p.bm0, p.bm1);
    }
    RT3D_LAUNCH_OK();
    // ---- B' (optional): statistical outlier removal of every frame's fused workspace (2cams_mask_cpu.py:530)
    const float *ws_src = ws_xyz;          // what the subtraction reads
    const int *ws_src_off = ws_off;
    int ws_src_stride = P.C;
    const int *ws_src_n = &p.cnt->ws_vox;
    if (h.ws_sor) {
        fb_ws_group_kernel<<<(P.F + 127) / 128, 128, 0, stream>>>(P, ws_off, p.ws_geom);
        fb_ws_rows_kernel<<<grid_for_n(h.ws_cap, kFbThreads, 8), kFbThreads, 0, stream>>>(P, ws_xyz, ws_off, &p.cnt->ws_vox, p.ws_pd, p.ws_grp);
        SorCore wc;
        wc.pd = p.ws_pd; wc.grp = p.ws_grp; wc.n_dev = &p.cnt->ws_vox; wc.cap = h.ws_cap; wc.geom = p.ws_geom; wc.ngroups = P.F;
        wc.total_bits = h.ws_sor_bits; wc.keys[0] = p.ws_keys[0]; wc.keys[1] = p.ws_keys[1]; wc.vals[0] = p.ws_vals[0]; wc.vals[1] = p.ws_vals[1];
        wc.sort_state = p.ws_sort_state; wc.ps = p.ws_ps; wc.avg = p.ws_avg; wc.col_tab = p.ws_coltab; wc.ticket = &p.cnt->ticket[8]; wc.status = &p.cnt->status;
        rc = launch_sor_core(P, wc, nullptr, stream);
        if (rc) return rc;
        rc = launch_sor_knn(P, wc, nullptr, stream);
        if (rc) return rc;
        FbWsSelectOp wop;
        wop.avg = p.ws_avg; wop.grp = p.ws_grp; wop.geom = p.ws_geom; wop.ws_xyz = ws_xyz; wop.filt = p.ws_filt; wop.filt_grp = p.ws_filt_grp;
        SelectCtl wcl;
        wcl.n = h.ws_cap; wcl.n_dev = &p.cnt->ws_vox; wcl.out_count = &p.cnt->ws_filt; wcl.state = p.ws_sel_state; wcl.ticket = &p.cnt->ticket[9];
        wcl.status = &p.cnt->status;
        RT3D_CUDA_OK(launch_select(wop, wcl, stream));
        fb_offsets_i32_kernel<<<(P.F + 256) / 256, 256, 0, stream>>>(p.ws_filt_grp, &p.cnt->ws_filt, 1, P.F, p.ws_filt_off);
        RT3D_LAUNCH_OK();
        ws_src = p.ws_filt; ws_src_off = p.ws_filt_off; ws_src_stride = 1; ws_src_n = &p.cnt->ws_filt;
    }
    RT3D_STAGE(2);
    const int gobj = grid_for_n(h.obj_cap, kFbThreads, 8);
    const bool centroid = P.fuse_mode == RT3D_FUSE_CENTROID;
    if (P.n_obj > 0) {
        // ---- C
        SortArgs sa;
        sa.keys[0] = p.keys[0]; sa.keys[1] = p.keys[1]; sa.vals[0] = sa.vals[1] = nullptr;
        sa.n = h.obj_cap; sa.n_dev = &p.cnt->obj_px; sa.total_bits = h.obj_total_bits; sa.total_bits_dev = nullptr;
        sa.state = p.sort_state[0]; sa.status = &p.cnt->status; sa.hist_done = false;
        rc = launch_sort(sa, stream);
        if (rc) return rc;
        FbObjUniqueOp uop;
        uop.keys = p.keys[sort_max_passes(h.obj_total_bits) & 1]; uop.P = P; uop.obj_pts = p.obj_pts; uop.obj_key = p.obj_key;
        SelectCtl c;
        c.n = h.obj_cap; c.n_dev = &p.cnt->obj_px; c.out_count = &p.cnt->obj_vox; c.state = p.sel_state[1];
        c.ticket = &p.cnt->ticket[1]; c.status = &p.cnt->status;
        RT3D_CUDA_OK(launch_select(uop, c, stream));
        fb_offsets_u64_kernel<<<(h.nseg + 256) / 256, 256, 0, stream>>>(p.obj_key, &p.cnt->obj_vox, P.seg_shift, h.nseg, p.seg_off);
        RT3D_STAGE(3);
        // ---- D
        fb_fuse_plan_kernel<<<(P.F + 127) / 128, 128, 0, stream>>>(P, p.seg_off, class_rank, p.geom, p.seg_dst, p.seg_group, p.cnt);
        fb_fuse_gather_kernel<<<gobj, kFbThreads, 0, stream>>>(P, p.obj_pts, p.obj_key, &p.cnt->obj_vox, p.seg_off, p.seg_dst, p.seg_group, p.pd, p.grp, p.gcode);
        SorCore sc;
        sc.pd = p.pd; sc.grp = p.grp; sc.n_dev = &p.cnt->obj_vox; sc.cap = h.obj_cap; sc.geom = p.geom; sc.ngroups = h.ngroups;
        sc.total_bits = h.sor_total_bits; sc.keys[0] = p.keys[0]; sc.keys[1] = p.keys[1]; sc.vals[0] = p.vals[0]; sc.vals[1] = p.vals[1];
        sc.sort_state = p.sort_state[1]; sc.ps = p.ps; sc.avg = p.avg; sc.col_tab = p.col_tab; sc.ticket = &p.cnt->ticket[5]; sc.status = &p.cnt->status;
        rc = launch_sor_core(P, sc, nullptr, stream);
        if (rc) return rc;
        const unsigned *svals = sc.svals;
        const int kblocks = persistent_blocks((h.obj_cap * 32 + kFbThreads - 1) / kFbThreads, 8);
        const unsigned char *need = nullptr;
        if (h.lat_words > 0) {   // bit planes first; what they cannot settle keeps need[j] = 1 for the grid search
            fb_lat_plan_kernel<<<h.ngroups, 128, 0, stream>>>(P, p.gcode, p.geom, p.lat, p.lat_cursor, (unsigned)h.lat_words);
            fb_lat_zero_kernel<<<kNumSMs * 8, kFbThreads, 0, stream>>>(p.lat_cursor, (unsigned)h.lat_words, reinterpret_cast<uint4 *>(p.lat_bits));
            fb_lat_fill_kernel<<<gobj, kFbThreads, 0, stream>>>(P, p.gcode, p.grp, &p.cnt->obj_vox, p.lat, p.lat_bits);
            fb_sor_lattice_kernel<<<grid_for_n(h.obj_cap, kLatThreads, 12), kLatThreads, 0, stream>>>(P, p.gcode, p.grp, svals, &p.cnt->obj_vox, p.geom, p.lat, p.lat_bits,
                                                                                                      p.avg, p.lat_need, &p.cnt->lat_settled);
            need = p.lat_need;
        }
        rc = launch_sor_knn(P, sc, need, stream);
        if (rc) return rc;
        FbSorSelectOp sop;
        sop.avg = p.avg; sop.pd = p.pd; sop.grp = p.grp; sop.geom = p.geom;
        sop.fused64 = centroid ? p.filt64 : obj_xyz; sop.fused32 = centroid ? nullptr : p.fused32; sop.fused_grp = p.fused_grp;
        c.n_dev = &p.cnt->obj_vox; c.out_count = &p.cnt->fused; c.state = p.sel_state[2]; c.ticket = &p.cnt->ticket[2];
        RT3D_CUDA_OK(launch_select(sop, c, stream));
        fb_offsets_i32_kernel<<<(h.ngroups + 256) / 256, 256, 0, stream>>>(p.fused_grp, &p.cnt->fused, 1, h.ngroups, centroid ? p.filt_off : obj_off);
        if (centroid) {
            fb_fuse_centroid_kernel<<<h.ngroups, 128, 0, stream>>>(p.geom, p.filt_off, p.filt64, p.seg_off, p.obj_pts, p.cen);
            fb_fuse_match_kernel<<<(P.F + 63) / 64, 64, 0, stream>>>(P, p.geom, p.filt_off, p.filt64, p.seg_off, p.obj_pts, p.cen, p.out, p.slot_n);
            fb_fuse_scan_kernel<<<1, kFbThreads, 0, stream>>>(p.slot_n, h.ngroups, (int)h.fused_cap, obj_off, p.cnt);
            fb_fuse_assemble_kernel<<<grid_for_n(h.fused_cap, kFbThreads, 8), kFbThreads, 0, stream>>>(p.out, obj_off, h.ngroups, p.filt64, p.obj_pts, obj_xyz,
                                                                                                      p.fused32, obj_meta);
        } else if (obj_meta != nullptr) {
            fb_meta_index_kernel<<<(h.ngroups + 255) / 256, 256, 0, stream>>>(P, p.geom, obj_off, obj_meta);
        }
        RT3D_LAUNCH_OK();
    }
    if (P.n_obj == 0) { RT3D_STAGE(3); }
    RT3D_STAGE(4);
    // ---- E
    if (P.n_obj > 0) {
        SubArgs sb;
        sb.F = P.F; sb.obj = p.fused32; sb.obj_off = obj_off; sb.off_stride = P.slots; sb.obj_cap = h.fused_cap;
        sb.ws = ws_src; sb.ws_off = ws_src_off; sb.C = ws_src_stride; sb.ws_cap = h.ws_cap; sb.ws_n_dev = ws_src_n; sb.t2 = P.sub_t2; sb.strict = 0;
        sb.keep = p.keep; sb.keys[0] = p.keys[0]; sb.keys[1] = p.keys[1]; sb.vals[0] = p.vals[0]; sb.vals[1] = p.vals[1];
        sb.sort_state = p.sort_state[2]; sb.grid = p.sgrid; sb.fine = p.fine; sb.leaf = p.leaf; sb.super = p.super; sb.fine_rep = p.fine_rep; sb.super_rep = p.super_rep; sb.obj_sorted = p.obj_sorted;
        sb.status = &p.cnt->status; sb.ticket = &p.cnt->ticket[6];
        sb.tile_cnt = reinterpret_cast<int *>(p.sel_state[3]); sb.tile_shift = kKeepTileShift;
        rc = launch_subtract(sb, stream);
        if (rc) return rc;
    } else {
        RT3D_CUDA_OK(cudaMemsetAsync(p.keep, 1, (size_t)h.ws_cap, stream));   // nothing to subtract: keep every row
    }
    FbKeepOp kop{ws_src, p.keep, kept_xyz};
    {
        const int ktiles = (int)((h.ws_cap + kKeepTile - 1) / kKeepTile);
        int *tile_off = reinterpret_cast<int *>(p.sel_state[3]);   // ktiles ints: counts, then offsets
        // with objects the flag kernel has counted the kept rows of every tile on the way
        if (P.n_obj == 0) fb_keep_count_kernel<<<ktiles, kFbThreads, 0, stream>>>(p.keep, ws_src_n, tile_off);
        fb_keep_scan_kernel<<<1, kKeepScanThreads, 0, stream>>>(tile_off, ws_src_n, &p.cnt->kept);
        fb_keep_write_kernel<<<ktiles, kFbThreads, 0, stream>>>(kop, ws_src_n, tile_off);
    }
    fb_frame_count_kernel<<<P.F, kFbThreads, 0, stream>>>(p.keep, ws_src_off, ws_src_stride, p.frame_cnt);
    fb_frame_scan_kernel<<<1, 32, 0, stream>>>(p.frame_cnt, P.F, kept_off);
    RT3D_LAUNCH_OK();
    // {status, obj_px, obj_vox, fused, ws_vox, kept, lat_settled, ws_filt}
    RT3D_CUDA_OK(cudaMemcpyAsync(counts, p.cnt, 8 * sizeof(int), cudaMemcpyDeviceToDevice, stream));
    RT3D_STAGE(5);
    if (timer) timer->recorded = true;
#undef RT3D_STAGE
    return RT3D_OK;
}
