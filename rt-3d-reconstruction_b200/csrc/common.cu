// common.cu -- host-side plumbing shared by all entry points of librt3d.so.
#include <stdarg.h>
#include <string.h>

#include "rt3d_common.cuh"

namespace rt3d {

static thread_local char g_last_error[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
}

}  // namespace rt3d

extern "C" int rt3d_abi_version(void) { return RT3D_ABI_VERSION; }
extern "C" const char *rt3d_last_error(void) { return rt3d::g_last_error; }

extern "C" size_t rt3d_scratch_bytes(int op, int64_t n, int64_t m) {
    if (n < 0) n = 0;
    if (m < 0) m = 0;
    switch (op) {
        case RT3D_OP_SELECT: return rt3d::select_scratch_bytes(n);
        case RT3D_OP_VOXEL_UNIQUE: return rt3d::voxel_scratch_bytes(n, false);
        case RT3D_OP_VOXEL_CENTROID: return rt3d::voxel_scratch_bytes(n, true);
        case RT3D_OP_SOR: return rt3d::sor_scratch_bytes(n);
        case RT3D_OP_SUBTRACT: return rt3d::subtract_scratch_bytes(n, m);
        case RT3D_OP_KEYS_UNIQUE: return rt3d::voxel_scratch_bytes(n, false);
        default: rt3d::set_error("rt3d_scratch_bytes: unknown op %d", op); return 0;
    }
}

// ------------------------------------------------------------------ verified fast division
namespace rt3d {

// every float with biased exponent in [27, 227) (i.e. 2^-100 .. 2^100), both signs, plus +-0
__global__ void verify_fast_div_kernel(DivBy fast, unsigned long long *mismatches) {
    const unsigned long long total = 2ull * 200ull * (1ull << 23);
    unsigned long long bad = 0;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < total;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned sign = (unsigned)(i & 1ull);
        const unsigned long long r = i >> 1;
        const unsigned bits = (sign << 31) | ((unsigned)(27 + (r >> 23)) << 23) | (unsigned)(r & 0x7fffffu);
        const float a = __uint_as_float(bits);
        const float want = __fdiv_rn(a, fast.b);
        const float got = fast(a);
        bad += (__float_as_uint(want) != __float_as_uint(got)) ? 1ull : 0ull;
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        // +0 must stay +0.  (-0 comes back as +0: no dividend of the path can be -0, and the two compare equal)
        if (__float_as_uint(fast(0.0f)) != __float_as_uint(__fdiv_rn(0.0f, fast.b))) ++bad;
    }
    if (bad) atomicAdd(mismatches, bad);
}

}  // namespace rt3d

#include <map>
#include <mutex>

namespace rt3d {

static std::mutex g_div_mutex;
static std::map<unsigned, bool> g_div_verified;   // key: bit pattern of fl32(b)

static unsigned divisor_key(double b) {
    const float f = (float)b;
    unsigned u;
    memcpy(&u, &f, sizeof u);
    return u;
}

bool fast_division_verified(double b) {
    std::lock_guard<std::mutex> lock(g_div_mutex);
    auto it = g_div_verified.find(divisor_key(b));
    return it != g_div_verified.end() && it->second;
}

static unsigned long long g_last_div_mismatches = 0;

bool verify_fast_division(double b, cudaStream_t stream) {
    {
        std::lock_guard<std::mutex> lock(g_div_mutex);
        auto it = g_div_verified.find(divisor_key(b));
        if (it != g_div_verified.end()) return it->second;
    }
    bool ok = false;
    const float bf = (float)b;
    if (bf > 1e-30f && bf < 1e30f) {
        unsigned long long *d_bad = nullptr;
        if (cudaMalloc(&d_bad, sizeof(unsigned long long)) == cudaSuccess) {
            cudaMemsetAsync(d_bad, 0, sizeof(unsigned long long), stream);
            verify_fast_div_kernel<<<kNumSMs * 16, 256, 0, stream>>>(DivBy::fast_ieee(b), d_bad);
            unsigned long long bad = 1;
            if (cudaMemcpyAsync(&bad, d_bad, sizeof bad, cudaMemcpyDeviceToHost, stream) == cudaSuccess &&
                cudaStreamSynchronize(stream) == cudaSuccess)
            { ok = (bad == 0); g_last_div_mismatches = bad; }
            cudaFree(d_bad);
        }
    }
    std::lock_guard<std::mutex> lock(g_div_mutex);
    g_div_verified[divisor_key(b)] = ok;
    return ok;
}

}  // namespace rt3d

// development aid: exhaustive mismatch count of the FMA-corrected division against __fdiv_rn for divisor b
extern "C" RT3D_API long long rt3d_debug_fast_division_mismatches(double b) {
    rt3d::verify_fast_division(b, 0);
    return (long long)rt3d::g_last_div_mismatches;
}
