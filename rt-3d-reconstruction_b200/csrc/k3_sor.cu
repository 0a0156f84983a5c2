// k3_sor.cu -- kernel family 3: statistical outlier removal.
//
// Replaces filter_outliers_sor (2cam/vision_pipeline_utils.py:48-58), i.e. Open3D 0.18
// PointCloud::RemoveStatisticalOutliers: a host-side nanoflann KD-tree + OpenMP k-NN loop in fp64.
// Here: points are binned into a uniform grid (cell keys radix-sorted with the point index as
// payload), one warp per query walks the neighbouring cell columns -- each column is one
// contiguous run of the sorted keys, found by binary search -- and keeps the k smallest squared
// distances in a warp-distributed sorted list maintained with ballots and shuffles.  All distance
// arithmetic is fp64 in nanoflann's operation order, sums run in Open3D's order (ascending per
// point, index order for the global mean / sigma), so the outlier mask is bit-exact against the
// CPU restatement (oracle/rt3d_oracle.c).
//
// Not HBM-bound: L2/shared-latency + fp64-ALU bound; reported as points/s.
#include <float.h>

#include "knn.cuh"
#include "radix_sort.cuh"
#include "select.cuh"

namespace rt3d {

constexpr int kSorThreads = 256;

__device__ __forceinline__ unsigned long long ordered_from_double(double x) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(x);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double double_from_ordered(unsigned long long u) {
    const unsigned long long b = (u >> 63) ? (u & 0x7fffffffffffffffull) : ~u;
    return __longlong_as_double((long long)b);
}

// Ctrl::dbl layout used by this file
enum { D_LO = 0 /*3*/, D_H = 3, D_INVH = 4, D_MU = 8, D_SIGMA = 9, D_THR = 10, D_VALID = 11 };
// Ctrl::layout: {dimx, dimy, dimz, bitsx, bitsy, bitsz, total_bits, 0}

// 1. fp64 working copy + bounding box
__global__ void __launch_bounds__(kSorThreads) sor_prepare_kernel(const void *__restrict__ xyz, int dtype, long long n,
                                                                  double *__restrict__ pd, unsigned long long *bbox) {
    double lo[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, hi[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
    for (long long i = blockIdx.x * (long long)kSorThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kSorThreads) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const double v = dtype == RT3D_F64 ? ((const double *)xyz)[3 * i + a] : (double)((const float *)xyz)[3 * i + a];
            pd[3 * i + a] = v;
            lo[a] = fmin(lo[a], v); hi[a] = fmax(hi[a], v);
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[a] = fmin(lo[a], __shfl_xor_sync(kFullMask, lo[a], o));
            hi[a] = fmax(hi[a], __shfl_xor_sync(kFullMask, hi[a], o));
        }
        if (lane_id() == 0 && lo[a] <= hi[a]) {
            atomicMin(&bbox[a], ordered_from_double(lo[a]));
            atomicMax(&bbox[3 + a], ordered_from_double(hi[a]));
        }
    }
}

__global__ void sor_init_bbox_kernel(unsigned long long *bbox) {
    if (threadIdx.x < 3) { bbox[threadIdx.x] = ~0ull; bbox[3 + threadIdx.x] = 0ull; }
}

// 2. grid geometry (1 thread).  Cell edge ~ 6 * sqrt(bbox face area / n): about 3-4 point
// spacings for the surface-like clouds of this pipeline, so the 27-cell neighbourhood usually
// holds the k = 20 neighbours.  Correctness never depends on this choice.
__global__ void sor_grid_kernel(const unsigned long long *bbox, long long n, Ctrl *ctrl) {
    double lo[3], ext[3];
    for (int a = 0; a < 3; ++a) {
        lo[a] = double_from_ordered(bbox[a]);
        ext[a] = double_from_ordered(bbox[3 + a]) - lo[a];
        if (!(ext[a] >= 0.0) || !isfinite(ext[a])) { ext[a] = 0.0; atomicOr(&ctrl->status, RT3D_DEV_KEY_RANGE); }
    }
    double e0 = ext[0], e1 = ext[1], e2 = ext[2], t;
    if (e0 < e1) { t = e0; e0 = e1; e1 = t; }
    if (e1 < e2) { t = e1; e1 = e2; e2 = t; }
    if (e0 < e1) { t = e0; e0 = e1; e1 = t; }
    const double area = e0 * (e1 > 0.0 ? e1 : e0);
    double h = 6.0 * sqrt(area / (double)(n > 0 ? n : 1));
    if (h < e0 / 1023.0) h = e0 / 1023.0;  // at most 1024 cells per axis (10 key bits)
    if (!(h > 0.0) || !isfinite(h)) h = 1.0;
    int total = 0;
    for (int a = 0; a < 3; ++a) {
        int dim = (int)floor(ext[a] / h) + 1;
        if (dim > 1024) dim = 1024;
        if (dim < 1) dim = 1;
        ctrl->layout[a] = dim;
        const int bits = dim <= 1 ? 0 : 32 - __clz((unsigned)(dim - 1));
        ctrl->layout[3 + a] = bits;
        total += bits;
        ctrl->dbl[D_LO + a] = lo[a];
    }
    ctrl->layout[6] = total;
    ctrl->total_bits = total;
    ctrl->dbl[D_H] = h;
    ctrl->dbl[D_INVH] = 1.0 / h;
}

// geometry of the single cloud handled by this file, read back from the control block
__device__ __forceinline__ void load_geom(GridGeom &g, const Ctrl *c) {
    for (int a = 0; a < 3; ++a) { g.lo[a] = c->dbl[D_LO + a]; g.dim[a] = c->layout[a]; g.bits[a] = c->layout[3 + a]; }
    g.h = c->dbl[D_H]; g.inv_h = c->dbl[D_INVH];
    g.prefix = 0ull;
}

// 3. cell keys (+ index payload, + digit histograms)
__global__ void __launch_bounds__(kSorThreads) sor_cellkey_kernel(const double *__restrict__ pd, long long n, const Ctrl *ctrl,
                                                                  unsigned long long *keys, unsigned *vals, unsigned *hist) {
    __shared__ unsigned s_hist[kRsMaxPasses * 256];
    GridGeom g;
    load_geom(g, ctrl);
    const int npass = (ctrl->total_bits + 7) / 8;
    for (int i = threadIdx.x; i < kRsMaxPasses * 256; i += kSorThreads) s_hist[i] = 0;
    __syncthreads();
    const long long stride = (long long)gridDim.x * kSorThreads;
    for (long long base = (long long)blockIdx.x * kSorThreads; base < n; base += stride) {
        const long long i = base + threadIdx.x;
        const bool valid = i < n;
        unsigned long long key = 0;
        if (valid) {
            key = g.pack(g.cell(pd[3 * i], 0), g.cell(pd[3 * i + 1], 1), g.cell(pd[3 * i + 2], 2));
            keys[i] = key;
            vals[i] = (unsigned)i;
        }
        rs_hist_add(s_hist, key, valid, npass);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < npass * 256; i += kSorThreads) {
        const unsigned c = s_hist[i];
        if (c) atomicAdd(&hist[i], c);
    }
}

// 4. points in cell order
__global__ void sor_gather_kernel(const double *__restrict__ pd, long long n, const unsigned *vals0, const unsigned *vals1,
                                  const Ctrl *ctrl, double *__restrict__ ps) {
    const unsigned *vals = (((ctrl->total_bits + 7) / 8) & 1) ? vals1 : vals0;
    for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < n; j += (long long)gridDim.x * blockDim.x) {
        const long long i = vals[j];
        ps[3 * j] = pd[3 * i]; ps[3 * j + 1] = pd[3 * i + 1]; ps[3 * j + 2] = pd[3 * i + 2];
    }
}

// 5. one warp per query (in cell order): mean distance to the k nearest neighbours, self included
template <int SLOTS>
__global__ void __launch_bounds__(kSorThreads) sor_knn_kernel(const double *__restrict__ ps, const unsigned long long *keys0,
                                                              const unsigned long long *keys1, const unsigned *vals0,
                                                              const unsigned *vals1, long long n, int k, const Ctrl *ctrl,
                                                              double *__restrict__ avg) {
    __shared__ KnnWarpScratch s_knn[kSorThreads / 32];
    const int sel = ((ctrl->total_bits + 7) / 8) & 1;
    const unsigned long long *keys = sel ? keys1 : keys0;
    const unsigned *vals = sel ? vals1 : vals0;
    GridGeom g;
    load_geom(g, ctrl);
    const int kk = (long long)k < n ? k : (int)n;
    // queries by atomic ticket (their cost varies widely), the next one requested while the current runs
    int *ticket = const_cast<int *>(&ctrl->ticket[1]);
    int tk = 0;
    if (lane_id() == 0) tk = atomicAdd(ticket, 1);
    for (;;) {
        const long long j = __shfl_sync(kFullMask, tk, 0);
        if (j >= n) break;
        if (lane_id() == 0) tk = atomicAdd(ticket, 1);
        int cx, cy, cz;
        g.unpack(keys[j], cx, cy, cz);
        const double m = knn_mean_distance<SLOTS>(g, keys, ps, 0, (int)n, ps[3 * j], ps[3 * j + 1], ps[3 * j + 2], cx, cy, cz, kk,
                                                  nullptr, &s_knn[threadIdx.x >> 5]);
        if (lane_id() == 0) avg[vals[j]] = m;
    }
}

// 6. global mean / sigma / threshold in Open3D's (sequential, index-order) summation order: 1 warp
__global__ void sor_stats_kernel(const double *__restrict__ avg, long long n, double std_ratio, Ctrl *ctrl, double *out_stats) {
    const int lane = lane_id();
    double s = 0.0;
    for (long long base = 0; base < n; base += 32) {
        const double v = (base + lane < n) ? avg[base + lane] : 0.0;
        const int m = (int)min((long long)32, n - base);
        for (int l = 0; l < m; ++l) {
            const double x = __shfl_sync(kFullMask, v, l);
            if (x > 0.0) s = __dadd_rn(s, x);
        }
    }
    const double valid = (double)n;  // every query finds at least itself
    const double mu = __ddiv_rn(s, valid);
    double sq = 0.0;
    for (long long base = 0; base < n; base += 32) {
        const double v = (base + lane < n) ? avg[base + lane] : 0.0;
        const int m = (int)min((long long)32, n - base);
        for (int l = 0; l < m; ++l) {
            const double x = __shfl_sync(kFullMask, v, l);
            const double dlt = __dsub_rn(x, mu);
            sq = __dadd_rn(sq, x > 0.0 ? __dmul_rn(dlt, dlt) : 0.0);
        }
    }
    const double sigma = sqrt(__ddiv_rn(sq, __dsub_rn(valid, 1.0)));
    const double thr = __dadd_rn(mu, __dmul_rn(std_ratio, sigma));
    if (lane == 0) {
        ctrl->dbl[D_MU] = mu; ctrl->dbl[D_SIGMA] = sigma; ctrl->dbl[D_THR] = thr; ctrl->dbl[D_VALID] = valid;
        if (out_stats) { out_stats[0] = mu; out_stats[1] = sigma; out_stats[2] = thr; out_stats[3] = valid; }
    }
}

// 7. keep mask + ordered compaction of the survivors (fp64 rows, input order)
struct SorSelectOp {
    using Payload = int;
    const double *avg;
    const double *pd;
    const Ctrl *ctrl;
    unsigned char *out_keep;
    double *out_xyz;
    __device__ __forceinline__ unsigned eval4(long long i0, long long n, int (&p)[4]) const {
        const double thr = ctrl->dbl[D_THR];
        unsigned keep = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const long long i = i0 + j;
            if (i >= n) break;
            const double a = avg[i];
            const bool k = a > 0.0 && a < thr;
            if (out_keep) out_keep[i] = k ? 1 : 0;
            if (k) { p[j] = (int)i; keep |= 1u << j; }
        }
        return keep;
    }
    __device__ __forceinline__ void store(long long pos, const int &i) const {
        if (!out_xyz) return;
        out_xyz[3 * pos] = pd[3 * (long long)i]; out_xyz[3 * pos + 1] = pd[3 * (long long)i + 1];
        out_xyz[3 * pos + 2] = pd[3 * (long long)i + 2];
    }
};

struct SorPlan {
    Ctrl *ctrl;
    unsigned long long *bbox;
    void *sort_state;
    unsigned long long *sel_state;
    size_t zero_bytes;
    double *pd, *ps, *avg;
    unsigned long long *keys[2];
    unsigned *vals[2];
    size_t total;
    bool ok;
};

static SorPlan plan_sor(void *scratch, size_t bytes, long long n) {
    Arena a(scratch, bytes);
    SorPlan p;
    p.ctrl = a.take<Ctrl>(1);
    p.bbox = a.take<unsigned long long>(8);
    p.sort_state = a.take<char>(sort_state_bytes(n, 32));
    p.sel_state = a.take<unsigned long long>(select_state_words(n));
    p.zero_bytes = a.off;
    p.pd = a.take<double>(3 * (size_t)n + 4);
    p.ps = a.take<double>(3 * (size_t)n + 4);
    p.avg = a.take<double>((size_t)n + 4);
    p.keys[0] = a.take<unsigned long long>((size_t)n + 4);
    p.keys[1] = a.take<unsigned long long>((size_t)n + 4);
    p.vals[0] = a.take<unsigned>((size_t)n + 4);
    p.vals[1] = a.take<unsigned>((size_t)n + 4);
    p.total = a.off;
    p.ok = a.ok();
    return p;
}

size_t sor_scratch_bytes(long long n) { return plan_sor(nullptr, 0, n).total; }

}  // namespace rt3d

using namespace rt3d;

extern "C" int rt3d_sor(const void *xyz, int dtype, int64_t n, int k, double std_ratio, uint8_t *out_keep, double *out_xyz,
                        int32_t *out_count, double *out_stats, double *out_avg, void *scratch, size_t scratch_bytes,
                        rt3d_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    RT3D_REQUIRE((xyz || n == 0) && out_count && scratch, RT3D_EINVAL, "rt3d_sor: bad arguments");
    RT3D_REQUIRE(dtype == RT3D_F32 || dtype == RT3D_F64, RT3D_EINVAL, "dtype must be RT3D_F32 or RT3D_F64");
    RT3D_REQUIRE(n >= 0 && n < (1ll << 30), RT3D_ERANGE, "point count %lld outside [0, 2^30)", (long long)n);
    // Open3D: "the number of neighbors and standard deviation ratio must be positive"
    RT3D_REQUIRE(k >= 1 && std_ratio > 0.0, RT3D_EINVAL,
                 "[RemoveStatisticalOutliers] Illegal input parameters, the number of neighbors and standard "
                 "deviation ratio must be positive.");
    RT3D_REQUIRE(k <= 64, RT3D_EINVAL, "rt3d_sor supports nb_neighbors <= 64 (got %d)", k);
    SorPlan p = plan_sor(scratch, scratch_bytes, n);
    RT3D_REQUIRE(p.ok, RT3D_ESCRATCH, "scratch too small: need %zu bytes, have %zu", p.total, scratch_bytes);
    RT3D_CUDA_OK(cudaMemsetAsync(scratch, 0, p.zero_bytes, stream));
    if (n == 0) { RT3D_CUDA_OK(cudaMemsetAsync(out_count, 0, sizeof(int), stream)); return RT3D_OK; }
    double *avg = out_avg ? out_avg : p.avg;
    const int gblocks = persistent_blocks((n + kSorThreads - 1) / kSorThreads, 8);

    sor_init_bbox_kernel<<<1, 32, 0, stream>>>(p.bbox);
    sor_prepare_kernel<<<gblocks, kSorThreads, 0, stream>>>(xyz, dtype, n, p.pd, p.bbox);
    sor_grid_kernel<<<1, 1, 0, stream>>>(p.bbox, n, p.ctrl);
    sor_cellkey_kernel<<<gblocks, kSorThreads, 0, stream>>>(p.pd, n, p.ctrl, p.keys[0], p.vals[0], sort_hist_ptr(p.sort_state));
    RT3D_LAUNCH_OK();
    SortArgs sa;
    sa.keys[0] = p.keys[0]; sa.keys[1] = p.keys[1]; sa.vals[0] = p.vals[0]; sa.vals[1] = p.vals[1];
    sa.n = n; sa.n_dev = nullptr; sa.total_bits = 30; sa.total_bits_dev = &p.ctrl->total_bits;
    sa.state = p.sort_state; sa.status = &p.ctrl->status; sa.hist_done = true;
    int rc = launch_sort(sa, stream);
    if (rc) return rc;
    sor_gather_kernel<<<gblocks, kSorThreads, 0, stream>>>(p.pd, n, p.vals[0], p.vals[1], p.ctrl, p.ps);
    // one warp per query
    const long long warps_needed = n;
    const int kblocks = persistent_blocks((warps_needed * 32 + kSorThreads - 1) / kSorThreads, 8);
    if (k <= 32)
        sor_knn_kernel<1><<<kblocks, kSorThreads, 0, stream>>>(p.ps, p.keys[0], p.keys[1], p.vals[0], p.vals[1], n, k, p.ctrl, avg);
    else
        sor_knn_kernel<2><<<kblocks, kSorThreads, 0, stream>>>(p.ps, p.keys[0], p.keys[1], p.vals[0], p.vals[1], n, k, p.ctrl, avg);
    sor_stats_kernel<<<1, 32, 0, stream>>>(avg, n, std_ratio, p.ctrl, out_stats);
    RT3D_LAUNCH_OK();
    SorSelectOp op{avg, p.pd, p.ctrl, out_keep, out_xyz};
    SelectCtl c;
    c.n = n; c.n_dev = nullptr; c.out_count = out_count; c.state = p.sel_state; c.ticket = &p.ctrl->ticket[0];
    c.status = &p.ctrl->status;
    RT3D_CUDA_OK(launch_select(op, c, stream));
    return RT3D_OK;
}
