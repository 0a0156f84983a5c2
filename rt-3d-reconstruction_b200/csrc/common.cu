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
