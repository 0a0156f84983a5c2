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

// Ctrl::dbl: {lo xyz, inv_h xyz}; Ctrl::layout: {dimx, dimy, dimz, ncells}
__global__ void sub_init_kernel(unsigned *bbox) {
    if (threadIdx.x < 3) { bbox[threadIdx.x] = 0xffffffffu; bbox[3 + threadIdx.x] = 0u; }
}

__global__ void __launch_bounds__(kSubThreads) sub_bbox_kernel(const float *__restrict__ obj, long long n, unsigned *bbox) {
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (long long i = blockIdx.x * (long long)kSubThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kSubThreads) {
#pragma unroll
        for (int a = 0; a < 3; ++a) { const float v = obj[3 * i + a]; lo[a] = fminf(lo[a], v); hi[a] = fmaxf(hi[a], v); }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[a] = fminf(lo[a], __shfl_xor_sync(kFullMask, lo[a], o));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(kFullMask, hi[a], o));
        }
        if (lane_id() == 0 && lo[a] <= hi[a]) {
            atomicMin(&bbox[a], ordered_from_float(lo[a]));
            atomicMax(&bbox[3 + a], ordered_from_float(hi[a]));
        }
    }
}

// coarse grid over the object cloud: cell edge >= thr / 4, at most 512 cells
__global__ void sub_grid_kernel(const unsigned *bbox, float thr, Ctrl *ctrl) {
    double lo[3], ext[3], emax = 0.0;
    for (int a = 0; a < 3; ++a) {
        lo[a] = (double)float_from_ordered(bbox[a]);
        ext[a] = (double)float_from_ordered(bbox[3 + a]) - lo[a];
        if (!(ext[a] >= 0.0) || !isfinite(ext[a])) ext[a] = 0.0;
        emax = fmax(emax, ext[a]);
    }
    double h = fmax((double)thr * 0.25, emax / 64.0);
    if (!(h > 0.0)) h = 1.0;
    int dim[3];
    for (;;) {
        long long cells = 1;
        for (int a = 0; a < 3; ++a) {
            dim[a] = (int)floor(ext[a] / h) + 1;
            cells *= dim[a];
        }
        if (cells <= kSubMaxCells) break;
        h *= 1.25;
    }
    for (int a = 0; a < 3; ++a) {
        ctrl->layout[a] = dim[a];
        ctrl->dbl[a] = lo[a];
        ctrl->dbl[3 + a] = 1.0 / h;
    }
    ctrl->layout[3] = dim[0] * dim[1] * dim[2];
}

__global__ void __launch_bounds__(kSubThreads) sub_cellkey_kernel(const float *__restrict__ obj, long long n, const Ctrl *ctrl,
                                                                  unsigned long long *keys, unsigned *vals, unsigned *hist) {
    __shared__ unsigned s_hist[kRsMaxPasses * 256];
    for (int i = threadIdx.x; i < kRsMaxPasses * 256; i += kSubThreads) s_hist[i] = 0;
    __syncthreads();
    const int dimy = ctrl->layout[1], dimz = ctrl->layout[2];
    const long long stride = (long long)gridDim.x * kSubThreads;
    for (long long base = (long long)blockIdx.x * kSubThreads; base < n; base += stride) {
        const long long i = base + threadIdx.x;
        const bool valid = i < n;
        unsigned long long key = 0;
        if (valid) {
            int c[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                c[a] = (int)floor(((double)obj[3 * i + a] - ctrl->dbl[a]) * ctrl->dbl[3 + a]);
                c[a] = c[a] < 0 ? 0 : (c[a] >= ctrl->layout[a] ? ctrl->layout[a] - 1 : c[a]);
            }
            key = (unsigned long long)((c[0] * dimy + c[1]) * dimz + c[2]);
            keys[i] = key;
            vals[i] = (unsigned)i;
        }
        rs_hist_add(s_hist, key, valid, 2);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * 256; i += kSubThreads) {
        const unsigned c = s_hist[i];
        if (c) atomicAdd(&hist[i], c);
    }
}

// one warp per coarse cell: its run in the sorted pairs, tight bounding box, points gathered in cell order
__global__ void __launch_bounds__(kSubThreads) sub_cells_kernel(const float *__restrict__ obj, long long n,
                                                                const unsigned long long *__restrict__ keys,
                                                                const unsigned *__restrict__ vals, const Ctrl *ctrl,
                                                                SubCell *cells, float *__restrict__ obj_sorted) {
    const int ncells = ctrl->layout[3];
    const int lane = lane_id();
    const int warp = (blockIdx.x * kSubThreads + threadIdx.x) >> 5;
    if (warp >= ncells) return;
    long long lo_i = 0, hi_i = n;
    {   // all lanes run the same two binary searches
        long long l = 0, h = n;
        while (l < h) { const long long m = (l + h) >> 1; if (keys[m] < (unsigned long long)warp) l = m + 1; else h = m; }
        lo_i = l;
        h = n;
        while (l < h) { const long long m = (l + h) >> 1; if (keys[m] <= (unsigned long long)warp) l = m + 1; else h = m; }
        hi_i = l;
    }
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (long long j = lo_i + lane; j < hi_i; j += 32) {
        const long long i = vals[j];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float v = obj[3 * i + a];
            obj_sorted[3 * j + a] = v;
            lo[a] = fminf(lo[a], v); hi[a] = fmaxf(hi[a], v);
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[a] = fminf(lo[a], __shfl_xor_sync(kFullMask, lo[a], o));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(kFullMask, hi[a], o));
        }
    }
    if (lane == 0) {
        SubCell c;
        for (int a = 0; a < 3; ++a) { c.lo[a] = lo[a]; c.hi[a] = hi[a]; }
        c.start = (int)lo_i; c.end = (int)hi_i;
        cells[warp] = c;
    }
}

// keep workspace row i iff every object point is farther than thr (exact predicate, fp64 at the margin)
struct SubtractOp {
    struct Payload { float x, y, z; };
    const float *ws;
    const float *obj_sorted;
    const SubCell *cells;
    const Ctrl *ctrl;
    float t2f;      // thr^2 in fp32
    double t2;      // (double)thr * (double)thr
    unsigned char *out_keep;
    float *out_xyz;

    __device__ __forceinline__ unsigned eval4(long long i0, long long n, Payload (&p)[4]) const {
        unsigned keep = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const long long i = i0 + j;
            if (i >= n) break;
            const float x = ws[3 * i], y = ws[3 * i + 1], z = ws[3 * i + 2];
            const bool k = !near_object(cells, ctrl->layout[3], obj_sorted, t2f, t2, x, y, z);
            if (out_keep) out_keep[i] = k ? 1 : 0;
            if (k) { p[j].x = x; p[j].y = y; p[j].z = z; keep |= 1u << j; }
        }
        return keep;
    }
    __device__ __forceinline__ void store(long long pos, const Payload &p) const {
        if (!out_xyz) return;
        float *o = out_xyz + 3 * pos;
        o[0] = p.x; o[1] = p.y; o[2] = p.z;
    }
};

struct SubPlan {
    Ctrl *ctrl;
    unsigned *bbox;
    void *sort_state;
    unsigned long long *sel_state;
    size_t zero_bytes;
    unsigned long long *keys[2];
    unsigned *vals[2];
    SubCell *cells;
    float *obj_sorted;
    size_t total;
    bool ok;
};

static SubPlan plan_sub(void *scratch, size_t bytes, long long m, long long n) {
    Arena a(scratch, bytes);
    SubPlan p;
    p.ctrl = a.take<Ctrl>(1);
    p.bbox = a.take<unsigned>(8);
    p.sort_state = a.take<char>(sort_state_bytes(n, 16));
    p.sel_state = a.take<unsigned long long>(select_state_words(m));
    p.zero_bytes = a.off;
    p.keys[0] = a.take<unsigned long long>((size_t)n + 4);
    p.keys[1] = a.take<unsigned long long>((size_t)n + 4);
    p.vals[0] = a.take<unsigned>((size_t)n + 4);
    p.vals[1] = a.take<unsigned>((size_t)n + 4);
    p.cells = a.take<SubCell>(kSubMaxCells);
    p.obj_sorted = a.take<float>(3 * (size_t)n + 4);
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

extern "C" int rt3d_subtract(const float *ws, int64_t m, const float *obj, int64_t n, float thr, uint8_t *out_keep,
                             float *out_xyz, int32_t *out_count, void *scratch, size_t scratch_bytes,
                             rt3d_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    RT3D_REQUIRE((ws || m == 0) && obj && out_count && scratch, RT3D_EINVAL, "rt3d_subtract: bad arguments");
    RT3D_REQUIRE(m >= 0 && m < (1ll << 30) && n < (1ll << 30), RT3D_ERANGE, "point counts outside [0, 2^30)");
    RT3D_REQUIRE(n >= 1, RT3D_EINVAL, "rt3d_subtract: empty object cloud (the reference raises IndexError here)");
    RT3D_REQUIRE(thr >= 0.0f, RT3D_EINVAL, "distance threshold must be non-negative");
    SubPlan p = plan_sub(scratch, scratch_bytes, m, n);
    RT3D_REQUIRE(p.ok, RT3D_ESCRATCH, "scratch too small: need %zu bytes, have %zu", p.total, scratch_bytes);
    RT3D_CUDA_OK(cudaMemsetAsync(scratch, 0, p.zero_bytes, stream));
    if (m == 0) { RT3D_CUDA_OK(cudaMemsetAsync(out_count, 0, sizeof(int), stream)); return RT3D_OK; }
    const int gblocks = persistent_blocks((n + kSubThreads - 1) / kSubThreads, 8);
    sub_init_kernel<<<1, 32, 0, stream>>>(p.bbox);
    sub_bbox_kernel<<<gblocks, kSubThreads, 0, stream>>>(obj, n, p.bbox);
    sub_grid_kernel<<<1, 1, 0, stream>>>(p.bbox, thr, p.ctrl);
    sub_cellkey_kernel<<<gblocks, kSubThreads, 0, stream>>>(obj, n, p.ctrl, p.keys[0], p.vals[0], sort_hist_ptr(p.sort_state));
    RT3D_LAUNCH_OK();
    SortArgs sa;
    sa.keys[0] = p.keys[0]; sa.keys[1] = p.keys[1]; sa.vals[0] = p.vals[0]; sa.vals[1] = p.vals[1];
    sa.n = n; sa.n_dev = nullptr; sa.total_bits = 16; sa.total_bits_dev = nullptr;  // cell ids < 512: two 8-bit passes
    sa.state = p.sort_state; sa.status = &p.ctrl->status; sa.hist_done = true;
    int rc = launch_sort(sa, stream);
    if (rc) return rc;
    // two passes -> the sorted pairs are back in buffer 0
    sub_cells_kernel<<<(kSubMaxCells * 32 + kSubThreads - 1) / kSubThreads, kSubThreads, 0, stream>>>(
        obj, n, p.keys[0], p.vals[0], p.ctrl, p.cells, p.obj_sorted);
    RT3D_LAUNCH_OK();
    SubtractOp op;
    op.ws = ws; op.obj_sorted = p.obj_sorted; op.cells = p.cells; op.ctrl = p.ctrl;
    op.t2 = (double)thr * (double)thr; op.t2f = (float)op.t2;
    op.out_keep = out_keep; op.out_xyz = out_xyz;
    SelectCtl c;
    c.n = m; c.n_dev = nullptr; c.out_count = out_count; c.state = p.sel_state; c.ticket = &p.ctrl->ticket[0];
    c.status = &p.ctrl->status;
    RT3D_CUDA_OK(launch_select(op, c, stream));
    return RT3D_OK;
}
