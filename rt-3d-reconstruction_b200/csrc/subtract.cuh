// subtract.cuh -- device side of the workspace subtraction shared by k4_fuse_subtract.cu and
// frame_batch.cu: the object cloud is binned into coarse cells with tight bounding boxes; a
// workspace point culls whole cells by box distance and scans only the ambiguous ones.
#pragma once

#include <float.h>

#include "rt3d_common.cuh"

namespace rt3d {

constexpr int kSubMaxCells = 512;

struct SubCell {  // 32 bytes
    float lo[3], hi[3];
    int start, end;
};

__device__ __forceinline__ unsigned ordered_from_float(float x) {
    const unsigned b = __float_as_uint(x);
    return (b >> 31) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float float_from_ordered(unsigned u) {
    return __uint_as_float((u >> 31) ? (u & 0x7fffffffu) : ~u);
}

// true iff some object point lies within thr of (x,y,z); exact at the margin (fp64 on the fp32 coordinates)
__device__ __forceinline__ bool near_object(const SubCell *__restrict__ cells, int ncells, const float *__restrict__ obj_sorted,
                                        float t2f, double t2, float x, float y, float z) {
    const float t2_hi = t2f * (1.0f + 1e-5f), t2_lo = t2f * (1.0f - 1e-5f);
    for (int c = 0; c < ncells; ++c) {
        const SubCell cell = cells[c];
        if (cell.start >= cell.end) continue;
        // squared distance from the point to the cell's tight box (min) and to its farthest corner (max)
        float dmin2 = 0.f, dmax2 = 0.f;
        const float q[3] = {x, y, z};
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float below = cell.lo[a] - q[a], above = q[a] - cell.hi[a];
            const float outside = fmaxf(fmaxf(below, above), 0.f);
            const float far = fmaxf(q[a] - cell.lo[a], cell.hi[a] - q[a]);
            dmin2 += outside * outside;
            dmax2 += far * far;
        }
        if (dmin2 > t2_hi) continue;           // surely farther than thr from everything in the cell
        if (dmax2 < t2_lo) return true;        // surely within thr of every point in the (non-empty) cell
        for (int j = cell.start; j < cell.end; ++j) {
            const float ox = obj_sorted[3 * j], oy = obj_sorted[3 * j + 1], oz = obj_sorted[3 * j + 2];
            const float dx = x - ox, dy = y - oy, dz = z - oz;
            const float d2f = dx * dx + dy * dy + dz * dz;
            if (d2f > t2f * (1.0f + 4e-6f)) continue;
            if (d2f < t2f * (1.0f - 4e-6f)) return true;
            // fp32 cannot decide: exact evaluation on the fp32 coordinates
            const double ex = (double)x - (double)ox, ey = (double)y - (double)oy, ez = (double)z - (double)oz;
            const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey)), __dmul_rn(ez, ez));
            if (d2 <= t2) return true;
        }
    }
    return false;
}


}  // namespace rt3d
