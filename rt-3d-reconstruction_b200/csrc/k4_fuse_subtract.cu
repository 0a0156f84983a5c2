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
            int c[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                c[k] = (int)floorf((a.obj[3 * i + k] - g.lo[k]) * g.inv_cell[k]);
                c[k] = c[k] < 0 ? 0 : (c[k] >= kSubFine ? kSubFine - 1 : c[k]);
            }
            key = ((unsigned long long)f << kSubCellBits) | (unsigned long long)((c[0] * kSubFine + c[1]) * kSubFine + c[2]);
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

// one warp per (frame, fine cell): its run in the sorted pairs, tight box, points gathered in cell order
__global__ void __launch_bounds__(kSubThreads) sub_cells_kernel(const SubArgs a, const unsigned long long *__restrict__ keys,
                                                                const unsigned *__restrict__ vals) {
    const int w = (blockIdx.x * kSubThreads + threadIdx.x) >> 5;
    const int f = w / kSubFineCells, cell = w - f * kSubFineCells;
    if (f >= a.F) return;
    const SubGrid g = a.grid[f];
    const int lane = lane_id();
    const unsigned long long key = ((unsigned long long)f << kSubCellBits) | (unsigned long long)cell;
    // the frame's points occupy [g.start, g.start + g.n) before and after the sort (frame id = high key bits)
    long long l = g.start, h = (long long)g.start + g.n;
    while (l < h) { const long long m = (l + h) >> 1; if (keys[m] < key) l = m + 1; else h = m; }
    const long long lo_i = l;
    h = (long long)g.start + g.n;
    while (l < h) { const long long m = (l + h) >> 1; if (keys[m] <= key) l = m + 1; else h = m; }
    const long long hi_i = l;
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (long long j = lo_i + lane; j < hi_i; j += 32) {
        const long long i = vals[j];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float v = a.obj[3 * i + k];
            a.obj_sorted[3 * j + k] = v;
            lo[k] = fminf(lo[k], v); hi[k] = fmaxf(hi[k], v);
        }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[k] = fminf(lo[k], __shfl_xor_sync(kFullMask, lo[k], o));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(kFullMask, hi[k], o));
        }
    }
    if (lane == 0) {
        SubCell c;
        for (int k = 0; k < 3; ++k) { c.lo[k] = lo[k]; c.hi[k] = hi[k]; }
        c.start = (int)lo_i; c.end = (int)hi_i;
        a.fine[(size_t)f * kSubFineCells + cell] = c;
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
}

// one thread per workspace row
__global__ void __launch_bounds__(kSubThreads) sub_flag_kernel(const SubArgs a, float t2f) {
    const int n = *a.ws_n_dev;
    const int nimg = a.F * a.C;
    for (long long i = blockIdx.x * (long long)kSubThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kSubThreads) {
        const int f = frame_of_row(a.ws_off, 1, nimg, (int)i) / a.C;
        const SubGrid g = a.grid[f];
        bool near = false;
        if (g.n > 0) {
            const float x = a.ws[3 * i], y = a.ws[3 * i + 1], z = a.ws[3 * i + 2];
            near = near_object(g, a.super + (size_t)f * kSubSuperCells, a.fine + (size_t)f * kSubFineCells, a.obj_sorted, t2f, a.t2, a.strict != 0, x, y, z);
        }
        a.keep[i] = near ? 0 : 1;
    }
}

int launch_subtract(const SubArgs &a, cudaStream_t stream) {
    RT3D_REQUIRE(a.F >= 1 && a.C >= 1 && a.t2 >= 0.0, RT3D_EINVAL, "subtract: bad arguments");
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
    SubCell *fine, *super;
    float *obj_sorted;
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
    p.super = a.take<SubCell>(kSubSuperCells);
    p.obj_sorted = a.take<float>(3 * (size_t)n + 4);
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
    a.sort_state = p.sort_state; a.grid = p.grid; a.fine = p.fine; a.super = p.super; a.obj_sorted = p.obj_sorted;
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
