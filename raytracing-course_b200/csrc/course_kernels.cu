// course_kernels.cu -- the two deterministic homework dialects (hw1 ray casting, hw2 Whitted), one thread per pixel.
// The per-pixel functions live in course_device.cuh.  hw3 and hw4 are Monte Carlo and run through the wavefront
// integrator of rt_kernels.cu (k_shade<HW3>).  Intersections go through the same Scene::RayIntersection as the hot
// path (planes + index BVH + replay); these scenes hold tens of primitives, so the frame is one launch and is bound
// by its own latency.
#include "course_device.cuh"
#include "rt_kernels.h"

namespace rtc {

__global__ void __launch_bounds__(128) k_raycast_hw1(DevScene S, float* accum) {
    const uint32_t npix = S.width * S.height;
    for (uint32_t pixel = blockIdx.x * blockDim.x + threadIdx.x; pixel < npix; pixel += gridDim.x * blockDim.x) {
        vec3 c = raycast_pixel_hw1(S, pixel);
        accum[3 * (size_t)pixel + 0] += c.x;
        accum[3 * (size_t)pixel + 1] += c.y;
        accum[3 * (size_t)pixel + 2] += c.z;
    }
}

__global__ void __launch_bounds__(128) k_whitted_hw2(DevScene S, float* accum) {
    const uint32_t npix = S.width * S.height;
    for (uint32_t pixel = blockIdx.x * blockDim.x + threadIdx.x; pixel < npix; pixel += gridDim.x * blockDim.x) {
        vec3 L = whitted_pixel_hw2(S, pixel);
        accum[3 * (size_t)pixel + 0] += L.x;
        accum[3 * (size_t)pixel + 1] += L.y;
        accum[3 * (size_t)pixel + 2] += L.z;
    }
}

__global__ void __launch_bounds__(256) k_resolve_flat(const float* accum, float inv_samples, uint32_t nvalues, uint8_t* out) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nvalues; i += gridDim.x * blockDim.x)
        out[i] = flat_u8(__fmul_rn(inv_samples, accum[i]));
}

static int grid_for_pixels(uint32_t n, int block, int sms) {
    uint32_t want = (n + block - 1) / block, cap = (uint32_t)sms * 8;
    if (want < 1) want = 1;
    return (int)(want < cap ? want : cap);
}
void launch_raycast_hw1(const LaunchCtx& c, const DevScene& S, float* accum) {
    k_raycast_hw1<<<grid_for_pixels(S.width * S.height, 128, c.sms), 128, 0, c.stream>>>(S, accum);
}
void launch_whitted_hw2(const LaunchCtx& c, const DevScene& S, float* accum) {
    k_whitted_hw2<<<grid_for_pixels(S.width * S.height, 128, c.sms), 128, 0, c.stream>>>(S, accum);
}
void launch_resolve_flat(const LaunchCtx& c, const float* accum, float inv_samples, uint32_t nvalues, uint8_t* out) {
    k_resolve_flat<<<grid_for_pixels(nvalues, 256, c.sms), 256, 0, c.stream>>>(accum, inv_samples, nvalues, out);
}

}  // namespace rtc
