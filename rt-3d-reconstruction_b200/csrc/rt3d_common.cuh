// rt3d_common.cuh -- shared host/device helpers of librt3d.so (sm_100a only).
//
// Everything on the path is HBM/L2-bound integer and fp32/fp64 scalar work: there
// are no tensor-core kernels here on purpose.  The building blocks are
//   * persistent grids sized from the 148 SMs, tiles handed out by an atomic ticket so that
//     decoupled look-back (single-pass ordered compaction / radix digit offsets) always waits
//     on a tile that is already resident;
//   * 128-bit read-only loads (ld.global.nc.L1::no_allocate) for streamed inputs;
//   * device-resident element counts, so a chain of stages never needs the host.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/rt3d.h"

namespace rt3d {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs
constexpr unsigned kFullMask = 0xffffffffu;

// ---------------------------------------------------------------- host-side errors
void set_error(const char *fmt, ...);

#define RT3D_REQUIRE(cond, code, ...)      \
    do {                                   \
        if (!(cond)) {                     \
            ::rt3d::set_error(__VA_ARGS__);\
            return (code);                 \
        }                                  \
    } while (0)

#define RT3D_CUDA_OK(expr)                                                                        \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            ::rt3d::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return RT3D_ECUDA;                                                                    \
        }                                                                                         \
    } while (0)

#define RT3D_LAUNCH_OK()                                                                          \
    do {                                                                                          \
        cudaError_t _e = cudaGetLastError();                                                      \
        if (_e != cudaSuccess) {                                                                  \
            ::rt3d::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
            return RT3D_ECUDA;                                                                    \
        }                                                                                         \
    } while (0)

// ---------------------------------------------------------------- scratch control block
// Every scratch arena starts with this block; the entry point zeroes it (one cudaMemsetAsync that
// also covers the look-back tile states that follow) before launching anything.
struct Ctrl {
    int status;        // RT3D_DEV_* bits
    int total_bits;    // radix sort: significant key bits (device-determined layouts)
    int count_a;       // stage-to-stage element counts
    int count_b;
    int ticket[28];    // one tile ticket per persistent kernel launch of the operation
    int layout[8];     // voxel key layout {min xyz, bits xyz, 0, 0}
    long long ll[4];   // bounding-box atomics etc.
    double dbl[16];    // grid parameters / statistics
};
static_assert(sizeof(Ctrl) % 16 == 0, "Ctrl must keep 16-byte alignment of what follows");

inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// carve a scratch arena: returns nullptr once the arena is exhausted
struct Arena {
    char *base;
    size_t size, off;
    Arena(void *p, size_t n) : base((char *)p), size(n), off(0) {}
    template <class T>
    T *take(size_t count) {
        size_t bytes = align_up(count * sizeof(T));
        if (base != nullptr && off + bytes > size) { off += bytes; return nullptr; }
        T *r = base ? (T *)(base + off) : nullptr;
        off += bytes;
        return r;
    }
    bool ok() const { return off <= size; }
};

// scratch sizing of each kernel family (defined next to the plan they mirror)
size_t select_scratch_bytes(long long n);
size_t voxel_scratch_bytes(long long n, bool with_index);
size_t sor_scratch_bytes(long long n);
size_t subtract_scratch_bytes(long long m, long long n);

inline int persistent_blocks(long long tiles, int per_sm) {
    long long cap = (long long)kNumSMs * per_sm;
    long long g = tiles < cap ? tiles : cap;
    return (int)(g < 1 ? 1 : g);
}

#ifdef __CUDACC__
// ---------------------------------------------------------------- device helpers
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// streamed, read-once inputs: 128-bit read-only loads that do not allocate in L1
__device__ __forceinline__ float4 ldg_stream_f4(const float *p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ unsigned ldg_stream_u32(const void *p) {
    unsigned r;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ float ldg_stream_f32(const float *p) {
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ longlong2 ldg_stream_ll2(const long long *p) {
    longlong2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.s64 {%0,%1}, [%2];" : "=l"(r.x), "=l"(r.y) : "l"(p));
    return r;
}

// gpu-scope relaxed accesses for look-back state words (value and status travel in one word,
// so no fence is needed around them)
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned *p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u32(unsigned *p, unsigned v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// look-back spins are bounded: a wedged predecessor becomes RT3D_DEV_LOOKBACK, never a hang
constexpr int kSpinLimit = 1 << 24;

constexpr unsigned long long kStAgg64 = 1ull << 62, kStPrefix64 = 2ull << 62, kStMask64 = 3ull << 62;
constexpr unsigned kStAgg32 = 1u << 30, kStPrefix32 = 2u << 30, kStMask32 = 3u << 30;

// warp-wide inclusive scan (int)
__device__ __forceinline__ int warp_incl_scan(int v) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(kFullMask, v, o);
        if (lane_id() >= o) v += t;
    }
    return v;
}
__device__ __forceinline__ long long warp_sum_ll(long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
    return v;
}

// Decoupled look-back over 64-bit tile states, run by one full warp.  Publishes this tile's
// aggregate, returns the exclusive prefix (sum of all earlier tiles) and publishes the inclusive
// prefix.  Tiles must have been handed out in ticket order.
__device__ __forceinline__ long long lookback_exclusive(unsigned long long *state, int tile, long long aggregate,
                                                        int *status) {
    const int lane = lane_id();
    if (tile == 0) {
        if (lane == 0) st_relaxed_u64(&state[0], kStPrefix64 | (unsigned long long)aggregate);
        return 0;
    }
    if (lane == 0) st_relaxed_u64(&state[tile], kStAgg64 | (unsigned long long)aggregate);
    long long excl = 0;
    int base = tile - 1;
    for (;;) {
        const int t = base - lane;
        unsigned long long w = kStPrefix64;  // tiles "before 0" carry an empty prefix
        if (t >= 0) {
            int spins = 0;
            do {
                w = ld_relaxed_u64(&state[t]);
            } while ((w & kStMask64) == 0 && ++spins < kSpinLimit);
            if ((w & kStMask64) == 0) {
                atomicOr(status, RT3D_DEV_LOOKBACK);
                w = kStPrefix64;
            }
        }
        const unsigned has_prefix = __ballot_sync(kFullMask, (w & kStMask64) == kStPrefix64);
        const int first = has_prefix ? (__ffs(has_prefix) - 1) : 32;
        long long v = (lane <= first) ? (long long)(w & ~kStMask64) : 0;
        excl += warp_sum_ll(v);
        if (has_prefix) break;
        base -= 32;
    }
    if (lane == 0) st_relaxed_u64(&state[tile], kStPrefix64 | (unsigned long long)(excl + aggregate));
    return excl;
}

// exclusive scan of one value per thread across a 256-thread block; `total` gets the block sum.
// s_tmp: 8 ints of shared memory; callers must __syncthreads() before reusing it.
__device__ __forceinline__ int block_excl_scan_256(int v, int *s_tmp, int &total) {
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    int incl = warp_incl_scan(v);
    if (lane == 31) s_tmp[warp] = incl;
    __syncthreads();
    int woff = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        int c = s_tmp[w];
        if (w < warp) woff += c;
        tot += c;
    }
    total = tot;
    return woff + incl - v;
}

// ---- the reference's fp32 arithmetic, spelled with round-to-nearest intrinsics so that nvcc can
// neither contract nor reassociate it (bit-exact voxel keys depend on it; SURVEY.md 7 hard part 1)
// `tensor / python_float`.  mode 0: IEEE division by fl32(b) (torch-CPU / NumPy).  mode 1: multiply by
// float(1.0 / double(b)) (torch-CUDA).  mode 2: IEEE division again, computed as one FMA-corrected
// multiply by the reciprocal (3 instructions instead of ~12); only selected after
// verify_fast_division() has compared it with __fdiv_rn for EVERY float dividend with an exponent in
// [-100, 100) and zero -- callers keep their dividends inside that range.
struct DivBy {
    float b, inv_b;
    int mode;
    __host__ __device__ DivBy() : b(1.f), inv_b(1.f), mode(0) {}
    __host__ DivBy(double b_, int flags) : b((float)b_), inv_b((float)(1.0 / b_)), mode(flags & 1) {}
    // the IEEE flavour with the fast path enabled (only after verification)
    __host__ static DivBy fast_ieee(double b_) {
        DivBy d;
        d.b = (float)b_; d.inv_b = (float)(1.0 / (double)d.b); d.mode = 2;
        return d;
    }
#ifdef __CUDACC__
    __device__ __forceinline__ float operator()(float a) const {
        if (mode == 2) {
            const float q = __fmul_rn(a, inv_b);
            const float r = __fmaf_rn(-q, b, a);
            return __fmaf_rn(r, inv_b, q);
        }
        return mode == 1 ? __fmul_rn(a, inv_b) : __fdiv_rn(a, b);
    }
#endif
};

// exhaustive check of DivBy mode 2 against __fdiv_rn for divisor fl32(b); cached per divisor.
// Launches a kernel and synchronises `stream`: call it from set-up paths only.
bool verify_fast_division(double b, cudaStream_t stream);
// cached result only (never launches): false when the divisor has not been verified
bool fast_division_verified(double b);

struct Rigid {  // p' = R p + t with torch.mm's rounding: fma chain over k, then a rounded +t
    float R[9], t[3];
    __device__ __forceinline__ void apply(float x, float y, float z, float &ox, float &oy, float &oz) const {
        float a0 = __fmul_rn(x, R[0]); a0 = __fmaf_rn(y, R[1], a0); a0 = __fmaf_rn(z, R[2], a0);
        float a1 = __fmul_rn(x, R[3]); a1 = __fmaf_rn(y, R[4], a1); a1 = __fmaf_rn(z, R[5], a1);
        float a2 = __fmul_rn(x, R[6]); a2 = __fmaf_rn(y, R[7], a2); a2 = __fmaf_rn(z, R[8], a2);
        ox = __fadd_rn(a0, t[0]); oy = __fadd_rn(a1, t[1]); oz = __fadd_rn(a2, t[2]);
    }
};

struct Box {  // closed intervals, crop_point_cloud_gpu (2cam/vision_pipeline_utils.py:78-84)
    float lo[3], hi[3];
    __device__ __forceinline__ bool contains(float x, float y, float z) const {
        return x >= lo[0] && x <= hi[0] && y >= lo[1] && y <= hi[1] && z >= lo[2] && z <= hi[2];
    }
};

__device__ __forceinline__ bool depth_valid(float d) {  // d > 0, not NaN, not inf (:26)
    return d > 0.0f && d <= 3.402823466e+38f;
}
#endif  // __CUDACC__

}  // namespace rt3d
