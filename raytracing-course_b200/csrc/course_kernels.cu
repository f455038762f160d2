// course_kernels.cu -- the two deterministic homework dialects, one thread per pixel (sm_100a):
//   hw1  ray casting: colour of the closest primitive              (hw1 src/scene.cpp:40-56, 167-179)
//   hw2  Whitted: ambient + point / directional lights with shadow rays, mirror reflection, dielectric
//        with the Schlick blend of BOTH branches                   (hw2 src/scene.cpp:262-341)
// hw3 and hw4 are Monte Carlo and run through the wavefront integrator of rt_kernels.cu (k_shade<HW3>).
// Intersections go through the same Scene::RayIntersection as the hot path (planes + index BVH + replay);
// these scenes hold tens of primitives, so the frame is one launch and is bound by its own latency.
#include "rt_device.cuh"
#include "rt_kernels.h"

namespace rtc {

RT_D SceneHit closest_hit(const DevScene& S, vec3 o, vec3 d) { return scene_intersect<0>(S, o, d, nullptr, nullptr, nullptr); }

// pixel-centre camera ray: Camera::get_to_ray(int x, int y), hw1 src/scene.cpp:30-38 / hw2 src/scene.cpp:229-237
RT_D void centre_ray(const DevScene& S, uint32_t pixel, vec3& o, vec3& d) {
    uint32_t x = pixel % S.width, y = pixel / S.width;
    camera_ray(S, (float)x + 0.5f, (float)y + 0.5f, o, d);
}

__global__ void __launch_bounds__(128) k_raycast_hw1(DevScene S, float* accum) {
    const uint32_t npix = S.width * S.height;
    for (uint32_t pixel = blockIdx.x * blockDim.x + threadIdx.x; pixel < npix; pixel += gridDim.x * blockDim.x) {
        vec3 o, d;
        centre_ray(S, pixel, o, d);
        SceneHit h = closest_hit(S, o, d);
        vec3 c = mk3(S.bg.x, S.bg.y, S.bg.z);
        if (h.id >= 0) c = ld3(ldg4(S.mat0 + h.id));
        accum[3 * (size_t)pixel + 0] += c.x;
        accum[3 * (size_t)pixel + 1] += c.y;
        accum[3 * (size_t)pixel + 2] += c.z;
    }
}

// DotLight::CalcLight / DirectedLight::CalcLight, hw2 src/lights.cpp:7-23
RT_D void calc_light(const DevScene& S, uint32_t l, vec3 p, vec3& colour, vec3& dir, float& dist) {
    float4 a = ldg4(S.plights + 4 * l);
    vec3 intensity = ld3(a);
    if (__float_as_uint(a.w)) {
        colour = intensity;
        dir = ld3(ldg4(S.plights + 4 * l + 3));
        dist = 1e18f;
        return;
    }
    vec3 to = ld3(ldg4(S.plights + 4 * l + 1)) - p;
    vec3 att = ld3(ldg4(S.plights + 4 * l + 2));
    dist = length(to);
    float k = 1.f / (att.x + att.y * dist + att.z * dist * dist);
    colour = k * intensity;
    dir = normalize(to);
}

// Scene::RayTrace of hw2 with the recursion turned into a stack of (ray, weight, remaining depth):
// L = sum over the leaves of the reflection / refraction tree of weight * local colour.
constexpr int kWhittedStack = 34;  // depth-first: at most ray_depth + 1 entries are pending
struct Pending {
    vec3 o, d, w;
    uint32_t depth;
};
__global__ void __launch_bounds__(128) k_whitted_hw2(DevScene S, float* accum) {
    const uint32_t npix = S.width * S.height;
    for (uint32_t pixel = blockIdx.x * blockDim.x + threadIdx.x; pixel < npix; pixel += gridDim.x * blockDim.x) {
        Pending st[kWhittedStack];
        int sp = 0;
        vec3 L = mk3(0, 0, 0);
        centre_ray(S, pixel, st[0].o, st[0].d);
        st[0].w = mk3(1, 1, 1);
        st[0].depth = S.ray_depth;
        sp = 1;
        while (sp > 0) {
            Pending r = st[--sp];
            if (r.depth == 0) continue;  // hw2 src/scene.cpp:263-265
            SceneHit h = closest_hit(S, r.o, r.d);
            if (h.id < 0) { L = L + r.w * mk3(S.bg.x, S.bg.y, S.bg.z); continue; }
            vec3 p = r.o + h.t * r.d;
            vec3 nd = normalize(r.d);
            float4 m0 = ldg4(S.mat0 + h.id);
            vec3 col = ld3(m0);
            uint32_t material = __float_as_uint(m0.w);
            vec3 rdir = reflect_dir(h.n, nd);
            if (material == MAT_DIFFUSE) {
                vec3 sum = mk3(S.ambient.x, S.ambient.y, S.ambient.z);
                for (uint32_t l = 0; l < S.nplights; ++l) {
                    vec3 lc, ldir;
                    float dist;
                    calc_light(S, l, p, lc, ldir, dist);
                    float k = dot(ldir, h.n);
                    if (k >= 0.f) {  // the light is not behind the surface
                        SceneHit b = closest_hit(S, p + S.eps * ldir, ldir);
                        if (!(b.id >= 0 && b.t <= dist)) sum = sum + k * lc;
                    }
                }
                L = L + r.w * (sum * col);
            } else if (material == MAT_METALLIC) {
                st[sp++] = Pending{p + S.eps * rdir, rdir, r.w * col, r.depth - 1};
            } else {  // DIELECTRIC, hw2 src/scene.cpp:300-330
                float eta1 = 1.f, eta2 = ldg4(S.mat1 + h.id).w;
                if (h.interior) { float tmp = eta1; eta1 = eta2; eta2 = tmp; }
                vec3 dir = -nd;
                float dn = dot(h.n, dir);
                float sin2 = eta1 / eta2 * sqrtf(1.f - dn * dn);
                if (fabsf(sin2) > 1.f) {  // total internal reflection
                    st[sp++] = Pending{p + S.eps * rdir, rdir, r.w, r.depth - 1};
                    continue;
                }
                float cos2 = sqrtf(1.f - sin2 * sin2);
                float e = eta1 / eta2;
                vec3 fr = e * (-dir) + (e * dn - cos2) * h.n;
                float q0 = (eta1 - eta2) / (eta1 + eta2);
                float r0 = q0 * q0;
                float m = 1.f - dn, m2 = m * m;
                float refl = r0 + (1.f - r0) * (m2 * m2 * m);
                vec3 wt = (1.f - refl) * r.w;
                if (!h.interior) wt = wt * col;
                st[sp++] = Pending{p + S.eps * fr, fr, wt, r.depth - 1};
                st[sp++] = Pending{p + S.eps * rdir, rdir, refl * r.w, r.depth - 1};
            }
        }
        accum[3 * (size_t)pixel + 0] += L.x;
        accum[3 * (size_t)pixel + 1] += L.y;
        accum[3 * (size_t)pixel + 2] += L.z;
    }
}

// hw1 writes the colour as it is: Color::toUInts, hw1 src/color.cpp:10-16 (no tone mapping, no gamma)
__global__ void __launch_bounds__(256) k_resolve_flat(const float* accum, float inv_samples, uint32_t nvalues, uint8_t* out) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nvalues; i += gridDim.x * blockDim.x)
        out[i] = (unsigned char)(int)roundf(__fmul_rn(255.f, __fmul_rn(inv_samples, accum[i])));
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
