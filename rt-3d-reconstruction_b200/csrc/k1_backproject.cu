// k1_backproject.cu -- kernel family 1: masked depth back-projection fused with the rigid
// transform and the workspace crop, plus the stand-alone forms of each step.
//
// Replaces (2cam/vision_pipeline_utils.py): torch.nonzero :296 (B0), convert_mask_to_3d_points
// :19-36 (B1), torch.mm(P,R.T)+t :301 / :239-240 (T1), crop_point_cloud_gpu :78-84 (C1) and the
// XYZ ingest of process_point_clouds :233-238 (W0) -- ~25 torch launches and 5 host syncs per
// object there, one ordered single-pass kernel here (select.cuh).
//
// HBM traffic per image (the roofline figure): 4*H*W (depth) + m*H*W (mask, m = 1 or
// 4 bytes; 0 when mask == NULL) + 12*N_out.
#include "select.cuh"

namespace rt3d {

struct Float3 { float x, y, z; };

__device__ __forceinline__ void store_xyz(float *out, long long pos, const Float3 &p) {
    float *o = out + 3 * pos;
    o[0] = p.x; o[1] = p.y; o[2] = p.z;
}

// ------------------------------------------------------------------------------------------
// depth (+mask) image -> points.  MASK: 0 = none (all pixels), 1 = uint8, 2 = float32.
struct Intrinsics {
    float cx, cy;
    DivBy div_fx, div_fy;
};

template <int MASK, bool VEC>
struct MaskBackprojectOp {
    using Payload = Float3;
    const float *depth;
    const void *mask;
    int W;
    Intrinsics k;
    bool has_rt, has_box;
    Rigid rt;
    Box box;
    float *out;

    __device__ __forceinline__ bool point(long long i, int u, int v, float d, bool m, Float3 &p) const {
        (void)i;
        if (!m || !depth_valid(d)) return false;
        float x = k.div_fx(__fmul_rn(__fsub_rn((float)u, k.cx), d));
        float y = k.div_fy(__fmul_rn(__fsub_rn((float)v, k.cy), d));
        float z = d;
        if (has_rt) rt.apply(x, y, z, x, y, z);
        if (has_box && !box.contains(x, y, z)) return false;
        p.x = x; p.y = y; p.z = z;
        return true;
    }

    __device__ __forceinline__ unsigned eval4(long long i0, long long n, Float3 (&p)[4]) const {
        unsigned keep = 0;
        if (VEC) {  // W % 4 == 0 and 16-byte aligned planes: the 4 items share an image row
            const float4 d4 = ldg_stream_f4(depth + i0);
            bool m[4] = {true, true, true, true};
            if (MASK == 1) {
                const unsigned w = ldg_stream_u32((const unsigned char *)mask + i0);
                m[0] = (w & 0xffu) != 0; m[1] = (w & 0xff00u) != 0; m[2] = (w & 0xff0000u) != 0; m[3] = (w >> 24) != 0;
            } else if (MASK == 2) {
                const float4 f = ldg_stream_f4((const float *)mask + i0);
                m[0] = f.x != 0.0f; m[1] = f.y != 0.0f; m[2] = f.z != 0.0f; m[3] = f.w != 0.0f;
            }
            const int v = (int)(i0 / W), u0 = (int)(i0 - (long long)v * W);
            const float d[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (point(i0 + j, u0 + j, v, d[j], m[j], p[j])) keep |= 1u << j;
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const long long i = i0 + j;
                if (i >= n) break;
                bool m = true;
                if (MASK == 1) m = ((const unsigned char *)mask)[i] != 0;
                if (MASK == 2) m = ((const float *)mask)[i] != 0.0f;
                const int v = (int)(i / W), u = (int)(i - (long long)v * W);
                if (point(i, u, v, depth[i], m, p[j])) keep |= 1u << j;
            }
        }
        return keep;
    }
    __device__ __forceinline__ void store(long long pos, const Float3 &p) const { store_xyz(out, pos, p); }
};

// ------------------------------------------------------------------------------------------
// B1 with the reference's exact signature: explicit (K,2) int64 [row, col] indices.
struct IndexBackprojectOp {
    using Payload = Float3;
    const long long *idx;
    const float *depth;
    int H, W;
    Intrinsics k;
    float *out;
    int *status;

    __device__ __forceinline__ unsigned eval4(long long i0, long long n, Float3 (&p)[4]) const {
        unsigned keep = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const long long i = i0 + j;
            if (i >= n) break;
            const longlong2 rc = ldg_stream_ll2(idx + 2 * i);
            long long v = rc.x, u = rc.y;
            if (v < 0) v += H;  // torch advanced indexing wraps negative indices
            if (u < 0) u += W;
            if (v < 0 || v >= H || u < 0 || u >= W) { atomicOr(status, RT3D_DEV_KEY_RANGE); continue; }
            const float d = __ldg(depth + v * W + u);
            if (!depth_valid(d)) continue;
            // `u_coords - cx` is int64 - python float -> fp32 in torch: fl32(u) - fl32(cx)
            p[j].x = k.div_fx(__fmul_rn(__fsub_rn((float)rc.y, k.cx), d));
            p[j].y = k.div_fy(__fmul_rn(__fsub_rn((float)rc.x, k.cy), d));
            p[j].z = d;
            keep |= 1u << j;
        }
        return keep;
    }
    __device__ __forceinline__ void store(long long pos, const Float3 &p) const { store_xyz(out, pos, p); }
};

// ------------------------------------------------------------------------------------------
// rows of an XYZ(+pad) image / an (N,3) cloud -> finite filter, transform, crop
template <int STRIDE>
struct XyzSelectOp {
    using Payload = Float3;
    const float *xyz;
    bool finite_filter, has_rt, has_box;
    Rigid rt;
    Box box;
    float *out;

    __device__ __forceinline__ unsigned eval4(long long i0, long long n, Float3 (&p)[4]) const {
        unsigned keep = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const long long i = i0 + j;
            if (i >= n) break;
            float x, y, z;
            if (STRIDE == 4) {
                const float4 q = ldg_stream_f4(xyz + 4 * i);
                x = q.x; y = q.y; z = q.z;
            } else {
                const float *q = xyz + (long long)STRIDE * i;
                x = ldg_stream_f32(q); y = ldg_stream_f32(q + 1); z = ldg_stream_f32(q + 2);
            }
            if (finite_filter && !(isfinite(x) && isfinite(y) && isfinite(z))) continue;
            if (has_rt) rt.apply(x, y, z, x, y, z);
            if (has_box && !box.contains(x, y, z)) continue;
            p[j].x = x; p[j].y = y; p[j].z = z;
            keep |= 1u << j;
        }
        return keep;
    }
    __device__ __forceinline__ void store(long long pos, const Float3 &p) const { store_xyz(out, pos, p); }
};

// ------------------------------------------------------------------------------------------
// B0: torch.nonzero(mask) -> (K,2) int64 [row, col]
template <int MASK>
struct NonzeroOp {
    using Payload = int2;
    const void *mask;
    int W;
    long long *out;
    __device__ __forceinline__ unsigned eval4(long long i0, long long n, int2 (&p)[4]) const {
        unsigned keep = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const long long i = i0 + j;
            if (i >= n) break;
            const bool m = MASK == 1 ? ((const unsigned char *)mask)[i] != 0 : ((const float *)mask)[i] != 0.0f;
            if (!m) continue;
            const int v = (int)(i / W);
            p[j] = make_int2(v, (int)(i - (long long)v * W));
            keep |= 1u << j;
        }
        return keep;
    }
    __device__ __forceinline__ void store(long long pos, const int2 &p) const {
        longlong2 rc; rc.x = p.x; rc.y = p.y;
        *reinterpret_cast<longlong2 *>(out + 2 * pos) = rc;
    }
};

// ------------------------------------------------------------------------------------------
__global__ void transform_kernel(const float *__restrict__ xyz, long long n, Rigid rt, float *__restrict__ out) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float x, y, z;
        rt.apply(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], x, y, z);
        out[3 * i] = x; out[3 * i + 1] = y; out[3 * i + 2] = z;
    }
}

struct Rigid64 { double R[9], t[3]; };
// T1d: double-precision fma chain (NumPy's dgemm order is BLAS-specific; 1e-12 tolerance)
__global__ void transform_f64_kernel(const float *__restrict__ xyz, long long n, Rigid64 rt, double *__restrict__ out) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            double a = __dmul_rn(x, rt.R[3 * j]);
            a = __fma_rn(y, rt.R[3 * j + 1], a);
            a = __fma_rn(z, rt.R[3 * j + 2], a);
            out[3 * i + j] = __dadd_rn(a, rt.t[j]);
        }
    }
}

// ------------------------------------------------------------------------------------------
static Intrinsics make_intrinsics(const double *intr, int flags) {
    Intrinsics k;
    k.cx = (float)intr[0]; k.cy = (float)intr[1];
    k.div_fx = DivBy(intr[2], flags);
    k.div_fy = DivBy(intr[3], flags);
    return k;
}
static void fill_rt(const float *Rt, Rigid &r) {
    for (int i = 0; i < 9; ++i) r.R[i] = Rt[i];
    for (int i = 0; i < 3; ++i) r.t[i] = Rt[9 + i];
}
static void fill_box(const float *b, Box &box) {
    for (int a = 0; a < 3; ++a) { box.lo[a] = b[2 * a]; box.hi[a] = b[2 * a + 1]; }
}

// common prologue: carve {Ctrl, look-back state} from scratch and zero it
static int select_prologue(long long n, int32_t *out_count, void *scratch, size_t scratch_bytes, cudaStream_t stream,
                           SelectCtl &c) {
    RT3D_REQUIRE(n >= 0 && n < (1ll << 30), RT3D_ERANGE, "item count %lld outside [0, 2^30)", n);
    RT3D_REQUIRE(out_count != nullptr && scratch != nullptr, RT3D_EINVAL, "out_count and scratch are required");
    Arena a(scratch, scratch_bytes);
    Ctrl *ctrl = a.take<Ctrl>(1);
    const size_t words = select_state_words(n);
    unsigned long long *state = a.take<unsigned long long>(words);
    RT3D_REQUIRE(a.ok(), RT3D_ESCRATCH, "scratch too small: need %zu bytes, have %zu", a.off, scratch_bytes);
    RT3D_CUDA_OK(cudaMemsetAsync(scratch, 0, a.off, stream));
    c.n = n; c.n_dev = nullptr; c.out_count = out_count; c.state = state;
    c.ticket = &ctrl->ticket[0]; c.status = &ctrl->status;
    return RT3D_OK;
}

size_t select_scratch_bytes(long long n) {
    Arena a(nullptr, 0);
    a.take<Ctrl>(1);
    a.take<unsigned long long>(select_state_words(n));
    return a.off;
}

}  // namespace rt3d

using namespace rt3d;

extern "C" int rt3d_mask_nonzero(const void *mask, int mask_dtype, int H, int W, int64_t *out_idx, int32_t *out_count,
                                 void *scratch, size_t scratch_bytes, rt3d_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    RT3D_REQUIRE(mask && out_idx && H >= 0 && W >= 0, RT3D_EINVAL, "rt3d_mask_nonzero: bad arguments");
    RT3D_REQUIRE(mask_dtype == RT3D_U8 || mask_dtype == RT3D_F32, RT3D_EINVAL, "mask dtype must be RT3D_U8 or RT3D_F32");
    SelectCtl c;
    int rc = select_prologue((long long)H * W, out_count, scratch, scratch_bytes, stream, c);
    if (rc) return rc;
    if (mask_dtype == RT3D_U8) {
        NonzeroOp<1> op{mask, W, (long long *)out_idx};
        RT3D_CUDA_OK(launch_select(op, c, stream));
    } else {
        NonzeroOp<2> op{mask, W, (long long *)out_idx};
        RT3D_CUDA_OK(launch_select(op, c, stream));
    }
    return RT3D_OK;
}

extern "C" int rt3d_backproject_indices(const int64_t *idx, int64_t K, const float *depth, int H, int W,
                                        const double *intr, int flags, float *out_xyz, int32_t *out_count, void *scratch,
                                        size_t scratch_bytes, rt3d_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    RT3D_REQUIRE((idx || K == 0) && depth && intr && out_xyz && H > 0 && W > 0, RT3D_EINVAL,
                 "rt3d_backproject_indices: bad arguments");
    RT3D_REQUIRE(((uintptr_t)idx & 15) == 0, RT3D_EINVAL, "index tensor must be 16-byte aligned");
    SelectCtl c;
    int rc = select_prologue(K, out_count, scratch, scratch_bytes, stream, c);
    if (rc) return rc;
    IndexBackprojectOp op{(const long long *)idx, depth, H, W, make_intrinsics(intr, flags), out_xyz, c.status};
    RT3D_CUDA_OK(launch_select(op, c, stream));
    return RT3D_OK;
}

template <int MASK, bool VEC>
static int launch_mask_bp(const float *depth, const void *mask, int W, const Intrinsics &k, const float *Rt,
                          const float *box, float *out, const SelectCtl &c, cudaStream_t stream) {
    MaskBackprojectOp<MASK, VEC> op;
    op.depth = depth; op.mask = mask; op.W = W; op.k = k; op.out = out;
    op.has_rt = Rt != nullptr; op.has_box = box != nullptr;
    if (Rt) fill_rt(Rt, op.rt);
    if (box) fill_box(box, op.box);
    RT3D_CUDA_OK(launch_select(op, c, stream));
    return RT3D_OK;
}

extern "C" int rt3d_backproject_mask(const float *depth, const void *mask, int mask_dtype, int H, int W,
                                     const double *intr, const float *Rt, const float *box, int flags, float *out_xyz,
                                     int32_t *out_count, void *scratch, size_t scratch_bytes, rt3d_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    RT3D_REQUIRE(depth && intr && out_xyz && H > 0 && W > 0, RT3D_EINVAL, "rt3d_backproject_mask: bad arguments");
    RT3D_REQUIRE(!mask || mask_dtype == RT3D_U8 || mask_dtype == RT3D_F32, RT3D_EINVAL,
                 "mask dtype must be RT3D_U8 or RT3D_F32");
    SelectCtl c;
    int rc = select_prologue((long long)H * W, out_count, scratch, scratch_bytes, stream, c);
    if (rc) return rc;
    const Intrinsics k = make_intrinsics(intr, flags);
    const int m = !mask ? 0 : (mask_dtype == RT3D_U8 ? 1 : 2);
    const bool vec = (W % 4 == 0) && (((uintptr_t)depth & 15) == 0) &&
                     (m == 0 || ((uintptr_t)mask & (m == 1 ? 3 : 15)) == 0);
#define RT3D_BP(MK)                                                                               \
    return vec ? launch_mask_bp<MK, true>(depth, mask, W, k, Rt, box, out_xyz, c, stream)          \
               : launch_mask_bp<MK, false>(depth, mask, W, k, Rt, box, out_xyz, c, stream)
    if (m == 0) { RT3D_BP(0); }
    if (m == 1) { RT3D_BP(1); }
    RT3D_BP(2);
#undef RT3D_BP
}

extern "C" int rt3d_xyz_select(const float *xyz, int64_t n, int stride, int finite_filter, const float *Rt,
                               const float *box, float *out_xyz, int32_t *out_count, void *scratch,
                               size_t scratch_bytes, rt3d_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    RT3D_REQUIRE((xyz || n == 0) && out_xyz && (stride == 3 || stride == 4), RT3D_EINVAL,
                 "rt3d_xyz_select: bad arguments (stride must be 3 or 4)");
    RT3D_REQUIRE(stride != 4 || ((uintptr_t)xyz & 15) == 0, RT3D_EINVAL, "stride-4 input must be 16-byte aligned");
    SelectCtl c;
    int rc = select_prologue(n, out_count, scratch, scratch_bytes, stream, c);
    if (rc) return rc;
    if (stride == 4) {
        XyzSelectOp<4> op;
        op.xyz = xyz; op.finite_filter = finite_filter != 0; op.has_rt = Rt != nullptr; op.has_box = box != nullptr;
        op.out = out_xyz;
        if (Rt) fill_rt(Rt, op.rt);
        if (box) fill_box(box, op.box);
        RT3D_CUDA_OK(launch_select(op, c, stream));
    } else {
        XyzSelectOp<3> op;
        op.xyz = xyz; op.finite_filter = finite_filter != 0; op.has_rt = Rt != nullptr; op.has_box = box != nullptr;
        op.out = out_xyz;
        if (Rt) fill_rt(Rt, op.rt);
        if (box) fill_box(box, op.box);
        RT3D_CUDA_OK(launch_select(op, c, stream));
    }
    return RT3D_OK;
}

extern "C" int rt3d_transform(const float *xyz, int64_t n, const float *Rt, float *out_xyz, rt3d_stream_t stream_) {
    RT3D_REQUIRE((xyz || n == 0) && Rt && (out_xyz || n == 0) && n >= 0, RT3D_EINVAL, "rt3d_transform: bad arguments");
    if (n == 0) return RT3D_OK;
    Rigid r;
    fill_rt(Rt, r);
    const int blocks = (int)((n + 255) / 256 < kNumSMs * 8 ? (n + 255) / 256 : kNumSMs * 8);
    transform_kernel<<<blocks, 256, 0, (cudaStream_t)stream_>>>(xyz, n, r, out_xyz);
    RT3D_LAUNCH_OK();
    return RT3D_OK;
}

extern "C" int rt3d_transform_f64(const float *xyz, int64_t n, const double *Rt, double *out_xyz, rt3d_stream_t stream_) {
    RT3D_REQUIRE((xyz || n == 0) && Rt && (out_xyz || n == 0) && n >= 0, RT3D_EINVAL, "rt3d_transform_f64: bad arguments");
    if (n == 0) return RT3D_OK;
    Rigid64 r;
    for (int i = 0; i < 9; ++i) r.R[i] = Rt[i];
    for (int i = 0; i < 3; ++i) r.t[i] = Rt[9 + i];
    const int blocks = (int)((n + 255) / 256 < kNumSMs * 8 ? (n + 255) / 256 : kNumSMs * 8);
    transform_f64_kernel<<<blocks, 256, 0, (cudaStream_t)stream_>>>(xyz, n, r, out_xyz);
    RT3D_LAUNCH_OK();
    return RT3D_OK;
}
