// k0_mask.cu -- mask pre-processing that sits directly in front of the back-projection.
//
//  erode_mask_gpu (2cam/vision_pipeline_utils.py:12-15): conv2d(mask, ones(k,k), padding=k//2) > 0, i.e. a
//      box *dilation*; output is (H + 1 - k%2... ) -- for even k the convolution grows the plane by one
//      row and column, exactly as torch's output shape does.
//  erode_mask (2cam/2cams_mask_cpu.py:54-57): cv2.erode with a k x k ones kernel (k = 12 there), anchor at
//      (k/2, k/2), border treated as "no constraint" (OpenCV's default +inf border for erosion).
//
// Both are window tests on a 0/1 plane; one thread per output pixel, rows of the window come from L1/L2.
#include "rt3d_common.cuh"

namespace rt3d {

template <class T>
__global__ void box_dilate_kernel(const T *__restrict__ mask, int H, int W, int k, float *__restrict__ out, int OH, int OW) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= OW || y >= OH) return;
    const int p = k / 2;
    float acc = 0.f;  // conv2d with a ones kernel and zero padding: the window sum
    for (int dy = 0; dy < k; ++dy) {
        const int yy = y - p + dy;
        if (yy < 0 || yy >= H) continue;
        for (int dx = 0; dx < k; ++dx) {
            const int xx = x - p + dx;
            if (xx < 0 || xx >= W) continue;
            acc += (float)mask[(size_t)yy * W + xx];
        }
    }
    out[(size_t)y * OW + x] = acc > 0.f ? 1.0f : 0.0f;
}

// cv2.erode, k x k rectangular kernel, anchor (k/2, k/2): out = min over the window, pixels outside the image ignored
__global__ void box_erode_kernel(const unsigned char *__restrict__ mask, int H, int W, int k, unsigned char *__restrict__ out) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    const int a = k / 2;
    unsigned char m = 255;
    for (int dy = 0; dy < k; ++dy) {
        const int yy = y - a + dy;
        if (yy < 0 || yy >= H) continue;
        for (int dx = 0; dx < k; ++dx) {
            const int xx = x - a + dx;
            if (xx < 0 || xx >= W) continue;
            const unsigned char v = mask[(size_t)yy * W + xx];
            m = v < m ? v : m;
        }
    }
    out[(size_t)y * W + x] = m;
}

}  // namespace rt3d

using namespace rt3d;

extern "C" int rt3d_box_dilate(const void *mask, int mask_dtype, int H, int W, int k, float *out, rt3d_stream_t stream_) {
    RT3D_REQUIRE(mask && out && H > 0 && W > 0 && k >= 1, RT3D_EINVAL, "rt3d_box_dilate: bad arguments");
    RT3D_REQUIRE(mask_dtype == RT3D_U8 || mask_dtype == RT3D_F32, RT3D_EINVAL, "mask dtype must be RT3D_U8 or RT3D_F32");
    const int OH = H + 2 * (k / 2) - k + 1, OW = W + 2 * (k / 2) - k + 1;
    dim3 b(32, 8), g((OW + 31) / 32, (OH + 7) / 8);
    if (mask_dtype == RT3D_U8) box_dilate_kernel<unsigned char><<<g, b, 0, (cudaStream_t)stream_>>>((const unsigned char *)mask, H, W, k, out, OH, OW);
    else box_dilate_kernel<float><<<g, b, 0, (cudaStream_t)stream_>>>((const float *)mask, H, W, k, out, OH, OW);
    RT3D_LAUNCH_OK();
    return RT3D_OK;
}

extern "C" int rt3d_box_erode(const uint8_t *mask, int H, int W, int k, int iterations, uint8_t *out, uint8_t *tmp,
                              rt3d_stream_t stream_) {
    RT3D_REQUIRE(mask && out && H > 0 && W > 0 && k >= 1 && iterations >= 1, RT3D_EINVAL, "rt3d_box_erode: bad arguments");
    RT3D_REQUIRE(iterations == 1 || tmp, RT3D_EINVAL, "rt3d_box_erode: iterations > 1 needs a tmp plane");
    dim3 b(32, 8), g((W + 31) / 32, (H + 7) / 8);
    const uint8_t *src = mask;
    for (int it = 0; it < iterations; ++it) {
        // ping-pong so that the last iteration lands in `out`
        uint8_t *dst = ((iterations - 1 - it) % 2 == 0) ? out : tmp;
        box_erode_kernel<<<g, b, 0, (cudaStream_t)stream_>>>(src, H, W, k, dst);
        src = dst;
    }
    RT3D_LAUNCH_OK();
    return RT3D_OK;
}
