// k2_voxel.cu -- kernel family 2: voxel-grid down-sampling on packed integer keys.
//
//  V1  downsample_point_cloud_gpu (2cam/vision_pipeline_utils.py:40-44): q = rint(p / v) * v, then
//      torch.unique(dim=0).  Here: voxel indices k = rint(p / v) are packed -- offset by the
//      per-axis minimum, with just enough bits per axis, x major -- into one integer whose order
//      IS the reference's lexicographic row order; keys are radix-sorted (radix_sort.cu), run
//      heads are compacted in order (select.cuh) and decoded back to k * v.
//  V2  downsample_point_cloud (2cam/2cams_mask_cpu.py:60-65, Open3D tensor voxel_down_sample):
//      key = floor(p / v); (key, point index) pairs are sorted, and one thread per voxel run
//      averages the fp32 positions in input order (segmented reduce).
//
// Algorithmic HBM bytes: 12*N in + 12*U out (V1); 12*N|24*N in + 12*U|24*U out (V2).
#include "radix_sort.cuh"
#include "select.cuh"

namespace rt3d {

constexpr int kKeyThreads = 256;

// --------------------------------------------------------------------------- voxel index
__device__ __forceinline__ bool voxel_index_round(float p, const DivBy &div_v, int &k) {
    const float q = rintf(div_v(p));                       // torch.round: half to even
    k = __float2int_rn(q);
    return fabsf(q) < 1.0e9f;                              // false for NaN / inf / absurd ranges
}

__device__ __forceinline__ int bits_for(unsigned range) { return range == 0 ? 0 : 32 - __clz(range); }

struct KeyLayout {
    int mn[3];
    int bits[3];
    __device__ __forceinline__ void load(const int *l) {
        mn[0] = l[0]; mn[1] = l[1]; mn[2] = l[2]; bits[0] = l[3]; bits[1] = l[4]; bits[2] = l[5];
    }
    __device__ __forceinline__ unsigned long long pack(int kx, int ky, int kz) const {
        return ((unsigned long long)(unsigned)(kx - mn[0]) << (bits[1] + bits[2])) |
               ((unsigned long long)(unsigned)(ky - mn[1]) << bits[2]) | (unsigned long long)(unsigned)(kz - mn[2]);
    }
    __device__ __forceinline__ void unpack(unsigned long long key, int &kx, int &ky, int &kz) const {
        kz = (int)(key & ((1ull << bits[2]) - 1ull)) + mn[2];
        ky = (int)((key >> bits[2]) & ((1ull << bits[1]) - 1ull)) + mn[1];
        kx = (int)(key >> (bits[1] + bits[2])) + mn[0];
    }
};

// pass A: per-axis min / max of the voxel indices (block reduce, then 6 atomics per block)
template <bool FLOOR64>
__global__ void __launch_bounds__(kKeyThreads) voxel_minmax_kernel(const void *__restrict__ xyz, int dtype, long long n_host,
                                                                   const int *n_dev, DivBy div_v, double voxel64,
                                                                   int *mm /* min xyz, max xyz */, int *status) {
    const long long n = n_dev ? (long long)*n_dev : n_host;
    int lo[3] = {INT_MAX, INT_MAX, INT_MAX}, hi[3] = {INT_MIN, INT_MIN, INT_MIN};
    bool bad = false;
    for (long long i = blockIdx.x * (long long)kKeyThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kKeyThreads) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            int k;
            bool ok;
            if (FLOOR64) {
                // Open3D: floor(p / v) in the input dtype
                double q;
                if (dtype == RT3D_F64) q = floor(__ddiv_rn(((const double *)xyz)[3 * i + a], voxel64));
                else q = (double)floorf(__fdiv_rn(((const float *)xyz)[3 * i + a], (float)voxel64));
                ok = fabs(q) < 1.0e9;
                k = __double2int_rn(q);
            } else {
                ok = voxel_index_round(((const float *)xyz)[3 * i + a], div_v, k);
            }
            if (!ok) { bad = true; continue; }
            lo[a] = min(lo[a], k); hi[a] = max(hi[a], k);
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[a] = min(lo[a], __shfl_xor_sync(kFullMask, lo[a], o));
            hi[a] = max(hi[a], __shfl_xor_sync(kFullMask, hi[a], o));
        }
    }
    if (lane_id() == 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            if (lo[a] <= hi[a]) { atomicMin(&mm[a], lo[a]); atomicMax(&mm[3 + a], hi[a]); }
        }
    }
    if (bad) atomicOr(status, RT3D_DEV_KEY_RANGE);
}

// pass B: 1 thread -- key layout from the index ranges
__global__ void voxel_layout_kernel(const int *mm, Ctrl *ctrl, int *out_layout) {
    int total = 0;
    int l[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int a = 0; a < 3; ++a) {
        const int lo = mm[a], hi = mm[3 + a];
        l[a] = lo <= hi ? lo : 0;
        l[3 + a] = lo <= hi ? bits_for((unsigned)(hi - lo)) : 0;
        total += l[3 + a];
    }
    if (total > 64) { atomicOr(&ctrl->status, RT3D_DEV_KEY_RANGE); total = 64; }
    l[6] = total;
    for (int i = 0; i < 8; ++i) { ctrl->layout[i] = l[i]; if (out_layout) out_layout[i] = l[i]; }
    ctrl->total_bits = total;
}

// pass C: keys (+ point index payload) and all digit histograms in one read of the points
template <bool FLOOR64, bool WITH_INDEX>
__global__ void __launch_bounds__(kKeyThreads) voxel_keygen_kernel(const void *__restrict__ xyz, int dtype, long long n_host,
                                                                   const int *n_dev, DivBy div_v, double voxel64,
                                                                   const Ctrl *ctrl, unsigned long long *keys,
                                                                   unsigned *vals, unsigned *hist) {
    __shared__ unsigned s_hist[kRsMaxPasses * 256];
    const long long n = n_dev ? (long long)*n_dev : n_host;
    KeyLayout L;
    L.load(ctrl->layout);
    const int npass = (ctrl->total_bits + 7) / 8;
    for (int i = threadIdx.x; i < kRsMaxPasses * 256; i += kKeyThreads) s_hist[i] = 0;
    __syncthreads();
    const long long stride = (long long)gridDim.x * kKeyThreads;
    for (long long base = (long long)blockIdx.x * kKeyThreads; base < n; base += stride) {
        const long long i = base + threadIdx.x;
        const bool valid = i < n;
        unsigned long long key = 0;
        if (valid) {
            int k[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                if (FLOOR64) {
                    double q;
                    if (dtype == RT3D_F64) q = floor(__ddiv_rn(((const double *)xyz)[3 * i + a], voxel64));
                    else q = (double)floorf(__fdiv_rn(((const float *)xyz)[3 * i + a], (float)voxel64));
                    k[a] = fabs(q) < 1.0e9 ? __double2int_rn(q) : L.mn[a];
                } else {
                    if (!voxel_index_round(((const float *)xyz)[3 * i + a], div_v, k[a])) k[a] = L.mn[a];
                }
            }
            key = L.pack(k[0], k[1], k[2]);
            keys[i] = key;
            if (WITH_INDEX) vals[i] = (unsigned)i;
        }
        rs_hist_add(s_hist, key, valid, npass);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < npass * 256; i += kKeyThreads) {
        const unsigned c = s_hist[i];
        if (c) atomicAdd(&hist[i], c);
    }
}

// --------------------------------------------------------------------------- V1 epilogue
// run heads of the sorted keys -> lattice points k * v, in key order (== torch.unique's row order)
struct UniqueDecodeOp {
    using Payload = unsigned long long;
    const unsigned long long *keys[2];
    const Ctrl *ctrl;
    float voxel;
    float *out_xyz;
    unsigned long long *out_keys;

    __device__ __forceinline__ const unsigned long long *src() const { return keys[((ctrl->total_bits + 7) / 8) & 1]; }
    __device__ __forceinline__ unsigned eval4(long long i0, long long n, unsigned long long (&p)[4]) const {
        const unsigned long long *k = src();
        unsigned keep = 0;
        unsigned long long prev = i0 > 0 ? k[i0 - 1] : 0ull;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const long long i = i0 + j;
            if (i >= n) break;
            const unsigned long long cur = k[i];
            if (i == 0 || cur != prev) { p[j] = cur; keep |= 1u << j; }
            prev = cur;
        }
        return keep;
    }
    __device__ __forceinline__ void store(long long pos, const unsigned long long &key) const {
        KeyLayout L;
        L.load(ctrl->layout);
        int kx, ky, kz;
        L.unpack(key, kx, ky, kz);
        float *o = out_xyz + 3 * pos;
        o[0] = __fmul_rn((float)kx, voxel); o[1] = __fmul_rn((float)ky, voxel); o[2] = __fmul_rn((float)kz, voxel);
        if (out_keys) out_keys[pos] = key;
    }
};

// --------------------------------------------------------------------------- V2 epilogue
// heads of voxel runs in the sorted (key, index) pairs -> compacted run start offsets
struct RunHeadOp {
    using Payload = int;
    const unsigned long long *keys[2];
    const Ctrl *ctrl;
    int *run_start;
    __device__ __forceinline__ const unsigned long long *src() const { return keys[((ctrl->total_bits + 7) / 8) & 1]; }
    __device__ __forceinline__ unsigned eval4(long long i0, long long n, int (&p)[4]) const {
        const unsigned long long *k = src();
        unsigned keep = 0;
        unsigned long long prev = i0 > 0 ? k[i0 - 1] : 0ull;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const long long i = i0 + j;
            if (i >= n) break;
            const unsigned long long cur = k[i];
            if (i == 0 || cur != prev) { p[j] = (int)i; keep |= 1u << j; }
            prev = cur;
        }
        return keep;
    }
    __device__ __forceinline__ void store(long long pos, const int &start) const { run_start[pos] = start; }
};

// one thread per voxel: mean of the member points' fp32 positions, accumulated in input order
// (the sort is stable, so the payload indices inside a run ascend)
__global__ void voxel_mean_kernel(const void *__restrict__ xyz, int dtype, long long n_host, const int *n_dev,
                                  const unsigned *vals0, const unsigned *vals1, const Ctrl *ctrl, const int *run_start,
                                  const int *num_runs, void *out_xyz, int *out_cnt) {
    const long long n = n_dev ? (long long)*n_dev : n_host;
    const unsigned *vals = (((ctrl->total_bits + 7) / 8) & 1) ? vals1 : vals0;
    const int U = *num_runs;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < U; r += gridDim.x * blockDim.x) {
        const int s = run_start[r];
        const int e = (r + 1 < U) ? run_start[r + 1] : (int)n;
        float acc[3] = {0.f, 0.f, 0.f};
        for (int j = s; j < e; ++j) {
            const long long i = vals[j];
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const float p = dtype == RT3D_F64 ? __double2float_rn(((const double *)xyz)[3 * i + a])
                                                  : ((const float *)xyz)[3 * i + a];
                acc[a] = __fadd_rn(acc[a], p);
            }
        }
        const float cnt = (float)(e - s);
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float m = __fdiv_rn(acc[a], cnt);
            if (dtype == RT3D_F64) ((double *)out_xyz)[3 * (long long)r + a] = (double)m;
            else ((float *)out_xyz)[3 * (long long)r + a] = m;
        }
        if (out_cnt) out_cnt[r] = e - s;
    }
}

// --------------------------------------------------------------------------- keys only
struct KeysUniqueOp {
    using Payload = unsigned long long;
    const unsigned long long *keys;
    unsigned long long *out_keys;
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
    __device__ __forceinline__ void store(long long pos, const unsigned long long &key) const { out_keys[pos] = key; }
};

__global__ void keys_decode_kernel(const unsigned long long *__restrict__ keys, long long n_host, const int *n_dev,
                                   const int *layout, float voxel, float *__restrict__ out) {
    const long long n = n_dev ? (long long)*n_dev : n_host;
    KeyLayout L;
    L.load(layout);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        int kx, ky, kz;
        L.unpack(keys[i], kx, ky, kz);
        out[3 * i] = __fmul_rn((float)kx, voxel); out[3 * i + 1] = __fmul_rn((float)ky, voxel);
        out[3 * i + 2] = __fmul_rn((float)kz, voxel);
    }
}

// keys packed with a per-call layout -> the canonical 3 x 21-bit biased layout (comparable across calls and GPUs)
__global__ void keys_rebase_kernel(const unsigned long long *__restrict__ keys, long long n_host, const int *n_dev,
                                   const int *layout, unsigned long long *__restrict__ out, int *status) {
    const long long n = n_dev ? (long long)*n_dev : n_host;
    KeyLayout L;
    L.load(layout);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        int k[3];
        L.unpack(keys[i], k[0], k[1], k[2]);
        unsigned long long o = 0;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const long long b = (long long)k[a] + (1ll << 20);
            if (b < 0 || b >= (1ll << 21)) atomicOr(status, RT3D_DEV_KEY_RANGE);
            o = (o << 21) | (unsigned long long)(b & ((1ll << 21) - 1));
        }
        out[i] = o;
    }
}

__global__ void init_minmax_kernel(int *mm) {
    if (threadIdx.x < 3) { mm[threadIdx.x] = INT_MAX; mm[3 + threadIdx.x] = INT_MIN; }
}

// --------------------------------------------------------------------------- scratch plans
struct VoxelPlan {
    Ctrl *ctrl;
    int *mm;
    unsigned long long *keys[2];
    unsigned *vals[2];
    void *sort_state;
    unsigned long long *sel_state;
    int *run_start;
    size_t zero_bytes;   // prefix of the arena that must be zeroed (ctrl, mm, sort state, select state)
    size_t total;
    bool ok;
};

static VoxelPlan plan_voxel(void *scratch, size_t bytes, long long n, bool with_index) {
    Arena a(scratch, bytes);
    VoxelPlan p;
    p.ctrl = a.take<Ctrl>(1);
    p.mm = a.take<int>(8);
    p.sort_state = a.take<char>(sort_state_bytes(n, 64));
    p.sel_state = a.take<unsigned long long>(select_state_words(n));
    p.zero_bytes = a.off;
    p.keys[0] = a.take<unsigned long long>((size_t)n + 4);
    p.keys[1] = a.take<unsigned long long>((size_t)n + 4);
    p.vals[0] = with_index ? a.take<unsigned>((size_t)n + 4) : nullptr;
    p.vals[1] = with_index ? a.take<unsigned>((size_t)n + 4) : nullptr;
    p.run_start = with_index ? a.take<int>((size_t)n + 4) : nullptr;
    p.total = a.off;
    p.ok = a.ok();
    return p;
}

size_t voxel_scratch_bytes(long long n, bool with_index) { return plan_voxel(nullptr, 0, n, with_index).total; }

static int grid_for(long long n, int threads, int per_sm) {
    return persistent_blocks((n + threads - 1) / threads, per_sm);
}

}  // namespace rt3d

using namespace rt3d;

extern "C" int rt3d_voxel_snap_unique(const float *xyz, int64_t n, const int32_t *n_dev, double voxel, int flags,
                                      float *out_xyz, uint64_t *out_keys, int32_t *out_layout, int32_t *out_count,
                                      void *scratch, size_t scratch_bytes, rt3d_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    RT3D_REQUIRE((xyz || n == 0) && out_xyz && out_count && scratch, RT3D_EINVAL, "rt3d_voxel_snap_unique: bad arguments");
    RT3D_REQUIRE(n >= 0 && n < (1ll << 30), RT3D_ERANGE, "point count %lld outside [0, 2^30)", (long long)n);
    RT3D_REQUIRE(voxel > 0.0, RT3D_EINVAL, "voxel size must be positive");
    VoxelPlan p = plan_voxel(scratch, scratch_bytes, n, false);
    RT3D_REQUIRE(p.ok, RT3D_ESCRATCH, "scratch too small: need %zu bytes, have %zu", p.total, scratch_bytes);
    RT3D_CUDA_OK(cudaMemsetAsync(scratch, 0, p.zero_bytes, stream));
    if (n == 0) { RT3D_CUDA_OK(cudaMemsetAsync(out_count, 0, sizeof(int), stream)); return RT3D_OK; }
    const DivBy div_v(voxel, flags);
    init_minmax_kernel<<<1, 32, 0, stream>>>(p.mm);
    voxel_minmax_kernel<false><<<grid_for(n, kKeyThreads, 8), kKeyThreads, 0, stream>>>(xyz, RT3D_F32, n, n_dev, div_v, 0.0,
                                                                                         p.mm, &p.ctrl->status);
    voxel_layout_kernel<<<1, 1, 0, stream>>>(p.mm, p.ctrl, out_layout);
    voxel_keygen_kernel<false, false><<<grid_for(n, kKeyThreads, 8), kKeyThreads, 0, stream>>>(
        xyz, RT3D_F32, n, n_dev, div_v, 0.0, p.ctrl, p.keys[0], nullptr, sort_hist_ptr(p.sort_state));
    RT3D_LAUNCH_OK();
    SortArgs sa;
    sa.keys[0] = p.keys[0]; sa.keys[1] = p.keys[1]; sa.vals[0] = sa.vals[1] = nullptr;
    sa.n = n; sa.n_dev = n_dev; sa.total_bits = 64; sa.total_bits_dev = &p.ctrl->total_bits;
    sa.state = p.sort_state; sa.status = &p.ctrl->status; sa.hist_done = true;
    int rc = launch_sort(sa, stream);
    if (rc) return rc;
    UniqueDecodeOp op;
    op.keys[0] = p.keys[0]; op.keys[1] = p.keys[1]; op.ctrl = p.ctrl; op.voxel = (float)voxel; op.out_xyz = out_xyz;
    op.out_keys = (unsigned long long *)out_keys;
    SelectCtl c;
    c.n = n; c.n_dev = n_dev; c.out_count = out_count; c.state = p.sel_state; c.ticket = &p.ctrl->ticket[0];
    c.status = &p.ctrl->status;
    RT3D_CUDA_OK(launch_select(op, c, stream));
    return RT3D_OK;
}

extern "C" int rt3d_voxel_centroid(const void *xyz, int dtype, int64_t n, double voxel, void *out_xyz, int32_t *out_cnt,
                                   int32_t *out_count, void *scratch, size_t scratch_bytes, rt3d_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    RT3D_REQUIRE((xyz || n == 0) && out_xyz && out_count && scratch, RT3D_EINVAL, "rt3d_voxel_centroid: bad arguments");
    RT3D_REQUIRE(dtype == RT3D_F32 || dtype == RT3D_F64, RT3D_EINVAL, "dtype must be RT3D_F32 or RT3D_F64");
    RT3D_REQUIRE(n >= 0 && n < (1ll << 30), RT3D_ERANGE, "point count %lld outside [0, 2^30)", (long long)n);
    RT3D_REQUIRE(voxel > 0.0, RT3D_EINVAL, "voxel size must be positive");
    VoxelPlan p = plan_voxel(scratch, scratch_bytes, n, true);
    RT3D_REQUIRE(p.ok, RT3D_ESCRATCH, "scratch too small: need %zu bytes, have %zu", p.total, scratch_bytes);
    RT3D_CUDA_OK(cudaMemsetAsync(scratch, 0, p.zero_bytes, stream));
    if (n == 0) { RT3D_CUDA_OK(cudaMemsetAsync(out_count, 0, sizeof(int), stream)); return RT3D_OK; }
    const DivBy none;
    init_minmax_kernel<<<1, 32, 0, stream>>>(p.mm);
    voxel_minmax_kernel<true><<<grid_for(n, kKeyThreads, 8), kKeyThreads, 0, stream>>>(xyz, dtype, n, nullptr, none, voxel,
                                                                                        p.mm, &p.ctrl->status);
    voxel_layout_kernel<<<1, 1, 0, stream>>>(p.mm, p.ctrl, nullptr);
    voxel_keygen_kernel<true, true><<<grid_for(n, kKeyThreads, 8), kKeyThreads, 0, stream>>>(
        xyz, dtype, n, nullptr, none, voxel, p.ctrl, p.keys[0], p.vals[0], sort_hist_ptr(p.sort_state));
    RT3D_LAUNCH_OK();
    SortArgs sa;
    sa.keys[0] = p.keys[0]; sa.keys[1] = p.keys[1]; sa.vals[0] = p.vals[0]; sa.vals[1] = p.vals[1];
    sa.n = n; sa.n_dev = nullptr; sa.total_bits = 64; sa.total_bits_dev = &p.ctrl->total_bits;
    sa.state = p.sort_state; sa.status = &p.ctrl->status; sa.hist_done = true;
    int rc = launch_sort(sa, stream);
    if (rc) return rc;
    RunHeadOp op;
    op.keys[0] = p.keys[0]; op.keys[1] = p.keys[1]; op.ctrl = p.ctrl; op.run_start = p.run_start;
    SelectCtl c;
    c.n = n; c.n_dev = nullptr; c.out_count = out_count; c.state = p.sel_state; c.ticket = &p.ctrl->ticket[0];
    c.status = &p.ctrl->status;
    RT3D_CUDA_OK(launch_select(op, c, stream));
    voxel_mean_kernel<<<grid_for(n, 128, 16), 128, 0, stream>>>(xyz, dtype, n, nullptr, p.vals[0], p.vals[1], p.ctrl,
                                                               p.run_start, out_count, out_xyz, out_cnt);
    RT3D_LAUNCH_OK();
    return RT3D_OK;
}

extern "C" int rt3d_keys_sort_unique(const uint64_t *keys, int64_t n, const int32_t *n_dev, int key_bits,
                                     uint64_t *out_keys, int32_t *out_count, void *scratch, size_t scratch_bytes,
                                     rt3d_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    RT3D_REQUIRE((keys || n == 0) && out_keys && out_count && scratch, RT3D_EINVAL, "rt3d_keys_sort_unique: bad arguments");
    RT3D_REQUIRE(n >= 0 && n < (1ll << 30), RT3D_ERANGE, "key count %lld outside [0, 2^30)", (long long)n);
    RT3D_REQUIRE(key_bits >= 1 && key_bits <= 64, RT3D_EINVAL, "key_bits must be in [1, 64]");
    VoxelPlan p = plan_voxel(scratch, scratch_bytes, n, false);
    RT3D_REQUIRE(p.ok, RT3D_ESCRATCH, "scratch too small: need %zu bytes, have %zu", p.total, scratch_bytes);
    RT3D_CUDA_OK(cudaMemsetAsync(scratch, 0, p.zero_bytes, stream));
    if (n == 0) { RT3D_CUDA_OK(cudaMemsetAsync(out_count, 0, sizeof(int), stream)); return RT3D_OK; }
    RT3D_CUDA_OK(cudaMemcpyAsync(p.keys[0], keys, (size_t)n * sizeof(uint64_t), cudaMemcpyDeviceToDevice, stream));
    SortArgs sa;
    sa.keys[0] = p.keys[0]; sa.keys[1] = p.keys[1]; sa.vals[0] = sa.vals[1] = nullptr;
    sa.n = n; sa.n_dev = n_dev; sa.total_bits = key_bits; sa.total_bits_dev = nullptr;
    sa.state = p.sort_state; sa.status = &p.ctrl->status; sa.hist_done = false;
    int rc = launch_sort(sa, stream);
    if (rc) return rc;
    KeysUniqueOp op;
    op.keys = p.keys[sort_max_passes(key_bits) & 1]; op.out_keys = (unsigned long long *)out_keys;
    SelectCtl c;
    c.n = n; c.n_dev = n_dev; c.out_count = out_count; c.state = p.sel_state; c.ticket = &p.ctrl->ticket[0];
    c.status = &p.ctrl->status;
    RT3D_CUDA_OK(launch_select(op, c, stream));
    return RT3D_OK;
}

extern "C" int rt3d_keys_decode(const uint64_t *keys, int64_t n, const int32_t *n_dev, const int32_t *layout, float voxel,
                                float *out_xyz, rt3d_stream_t stream_) {
    RT3D_REQUIRE((keys || n == 0) && layout && (out_xyz || n == 0) && n >= 0, RT3D_EINVAL, "rt3d_keys_decode: bad arguments");
    if (n == 0) return RT3D_OK;
    keys_decode_kernel<<<grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream_>>>((const unsigned long long *)keys, n, n_dev,
                                                                                layout, voxel, out_xyz);
    RT3D_LAUNCH_OK();
    return RT3D_OK;
}

extern "C" int rt3d_keys_rebase(const uint64_t *keys, int64_t n, const int32_t *n_dev, const int32_t *layout,
                                uint64_t *out_keys, int32_t *out_layout, int32_t *status, rt3d_stream_t stream_) {
    RT3D_REQUIRE((keys || n == 0) && layout && (out_keys || n == 0) && status && n >= 0, RT3D_EINVAL, "rt3d_keys_rebase: bad arguments");
    cudaStream_t stream = (cudaStream_t)stream_;
    if (out_layout) {   // canonical layout: min = -2^20 per axis, 21 bits per axis
        const int canon[8] = {-(1 << 20), -(1 << 20), -(1 << 20), 21, 21, 21, 63, 0};
        RT3D_CUDA_OK(cudaMemcpyAsync(out_layout, canon, sizeof canon, cudaMemcpyHostToDevice, stream));
    }
    if (n == 0) return RT3D_OK;
    keys_rebase_kernel<<<grid_for(n, 256, 8), 256, 0, stream>>>((const unsigned long long *)keys, n, n_dev, layout,
                                                               (unsigned long long *)out_keys, status);
    RT3D_LAUNCH_OK();
    return RT3D_OK;
}
