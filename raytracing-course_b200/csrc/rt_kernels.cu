// rt_kernels.cu -- kernels of the wavefront integrator (generate / extend / shade / resolve)
// and the batch "probe" kernels behind rtc_intersect, rtc_mix_pdf, ... (sm_100a).
#include "rt_device.cuh"
#include "rt_kernels.h"

namespace rtc {

// ------------------------------------------------------------------------------- wavefront
// Queue header words in device memory (one set per batch, zeroed before the batch):
//   q[0]            number of camera paths generated
//   q[b]  (b>=1)    number of paths alive after the shading of bounce b (input of extend b+1)
// Every kernel reads its element count from there: no host round trip between bounces.

// Scene::Sample's jitter + Camera::GetToRay (src/scene.cpp:189-200): one thread per path.
// Path ids run sample-major over the image: consecutive threads = consecutive pixels of a row.
__global__ void __launch_bounds__(256) k_generate(DevScene S, PathSoA P, uint32_t* q, uint64_t first_path, uint32_t count,
                                                   uint32_t seed, uint32_t sample_begin) {
    const uint32_t npix = S.width * S.height;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
        uint64_t pid = first_path + i;
        uint32_t sample = sample_begin + (uint32_t)(pid / npix);
        uint32_t pixel = (uint32_t)(pid % npix);
        uint32_t x = pixel % S.width, y = pixel / S.width;
        Rng g{seed, pixel, sample, 0};
        uint4 b = g.block(0);
        float fx = __fadd_rn((float)x, u01(b.x)), fy = __fadd_rn((float)y, u01(b.y));
        vec3 o, d;
        camera_ray(S, fx, fy, o, d);
        P.o[i] = make_float4(o.x, o.y, o.z, 0.f);
        P.d[i] = make_float4(d.x, d.y, d.z, 0.f);
        P.beta[i] = make_float4(1.f, 1.f, 1.f, __uint_as_float(pixel));
        P.rad[i] = make_float4(0.f, 0.f, 0.f, __uint_as_float(sample));
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) q[0] = count;
}

// Scene::RayIntersection for every queued ray.
template <int MODE, bool STATS>
__global__ void __launch_bounds__(128) k_extend(DevScene S, PathSoA P, HitSoA H, const uint32_t* qcount,
                                                 unsigned long long* stats) {
    const uint32_t count = *qcount;
    uint32_t visits = 0, tests = 0, fallbacks = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
        float4 o4 = P.o[i], d4 = P.d[i];
        vec3 o = ld3(o4), d = ld3(d4);
        SceneHit h = scene_intersect<MODE>(S, o, d, STATS ? &visits : nullptr, STATS ? &tests : nullptr, &fallbacks);
        H.tn[i] = make_float4(h.t, h.n.x, h.n.y, h.n.z);
        H.id[i] = h.id < 0 ? HIT_MISS : ((uint32_t)h.id | (h.interior ? HIT_INTERIOR : 0u));
    }
    if (fallbacks) atomicAdd(stats + 5, (unsigned long long)fallbacks);
    if (STATS) {
        if (visits) atomicAdd(stats + 4, (unsigned long long)visits);
        if (tests) atomicAdd(stats + 6, (unsigned long long)tests);
    }
}

RT_D void deposit(float* accum, uint32_t pixel, vec3 L) {
    atomicAdd(accum + 3 * (size_t)pixel + 0, L.x);
    atomicAdd(accum + 3 * (size_t)pixel + 1, L.y);
    atomicAdd(accum + 3 * (size_t)pixel + 2, L.z);
}

// Scene::RayTrace's material switch (src/scene.cpp:96-177) for one bounce, recursion unrolled:
// L = sum_k beta_k * E_k.  Surviving paths are written compacted (warp ballot + one atomic per
// warp) into the next queue; finished paths add their radiance to the pixel sum.
__global__ void __launch_bounds__(256) k_shade(DevScene S, PathSoA P, HitSoA H, PathSoA N, const uint32_t* qin, uint32_t* qout,
                                                float* accum, uint32_t bounce, uint32_t seed) {
    const uint32_t count = *qin;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t rounded = (count + 31u) & ~31u;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < rounded; i += gridDim.x * blockDim.x) {
        bool alive = false;
        vec3 no = mk3(0, 0, 0), nd = mk3(0, 0, 0), beta = mk3(0, 0, 0), L = mk3(0, 0, 0);
        uint32_t pixel = 0, sample = 0;
        if (i < count) {
            float4 b4 = P.beta[i], r4 = P.rad[i];
            beta = ld3(b4); L = ld3(r4);
            pixel = __float_as_uint(b4.w); sample = __float_as_uint(r4.w);
            uint32_t hid = H.id[i];
            if (hid == HIT_MISS) {
                L = L + beta * mk3(S.bg.x, S.bg.y, S.bg.z);  // src/scene.cpp:92-94
            } else {
                float4 tn = H.tn[i];
                uint32_t prim = hid & 0xFFFFFFu;
                bool interior = (hid & HIT_INTERIOR) != 0;
                vec3 normal = mk3(tn.y, tn.z, tn.w);
                vec3 o = ld3(P.o[i]), d = ld3(P.d[i]);
                vec3 p = o + tn.x * d;
                float4 m0 = ldg4(S.mat0 + prim), m1 = ldg4(S.mat1 + prim);
                vec3 col = ld3(m0);
                L = L + beta * ld3(m1);
                uint32_t material = __float_as_uint(m0.w);
                if (bounce < S.ray_depth) {
                    Rng g{seed, pixel, sample, bounce};
                    if (material == MAT_DIFFUSE) {
                        vec3 p_outer = p + kSceneEps * normal;
                        vec3 dir = mix_sample(S, g, p_outer, normal);
                        float cs = dot(dir, normal);
                        if (cs > 0.f) {
                            float pw = mix_pdf(S, p_outer, normal, dir);
                            vec3 w = mk3(col.x / kPi, col.y / kPi, col.z / kPi);
                            float k2 = 1.f / pw;
                            beta = beta * mk3(w.x * cs * k2, w.y * cs * k2, w.z * cs * k2);
                            no = p + kSceneEps * dir; nd = dir;
                            alive = true;
                        }
                    } else if (material == MAT_METALLIC) {
                        vec3 rd = reflect_dir(normal, normalize(d));
                        beta = beta * col;
                        no = p + kSceneEps * rd; nd = rd;
                        alive = true;
                    } else {  // DIELECTRIC, src/scene.cpp:130-170
                        float eta1 = 1.f, eta2 = m1.w;
                        if (interior) { float tmp = eta1; eta1 = eta2; eta2 = tmp; }
                        vec3 nd_in = normalize(d);
                        vec3 dir = -nd_in;
                        float dn = dot(normal, dir);
                        float sin2 = eta1 / eta2 * sqrtf(fmaxf(0.f, 1.f - dn * dn));
                        bool reflect = fabsf(sin2) > 1.f;
                        if (!reflect) {
                            float q0 = (eta1 - eta2) / (eta1 + eta2);
                            float r0 = q0 * q0;
                            float m = 1.f - dn;
                            float m2 = m * m;
                            float r = r0 + (1.f - r0) * (m2 * m2 * m);
                            reflect = u01(g.block(0).x) < r;
                        }
                        if (reflect) {
                            vec3 rd = reflect_dir(normal, nd_in);
                            no = p + kSceneEps * rd; nd = rd;
                        } else {
                            float cos2 = sqrtf(1.f - sin2 * sin2);
                            float e = eta1 / eta2;
                            vec3 fr = e * (-dir) + (e * dn - cos2) * normal;
                            no = p + kSceneEps * fr; nd = fr;
                            if (!interior) beta = beta * col;
                        }
                        alive = true;
                    }
                }
            }
            if (!alive) deposit(accum, pixel, L);
        }
        unsigned mask = __ballot_sync(0xFFFFFFFFu, alive);
        if (mask) {
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(qout, (uint32_t)__popc(mask));
            base = __shfl_sync(0xFFFFFFFFu, base, 0);
            if (alive) {
                uint32_t dst = base + __popc(mask & ((1u << lane) - 1u));
                N.o[dst] = make_float4(no.x, no.y, no.z, 0.f);
                N.d[dst] = make_float4(nd.x, nd.y, nd.z, 0.f);
                N.beta[dst] = make_float4(beta.x, beta.y, beta.z, __uint_as_float(pixel));
                N.rad[dst] = make_float4(L.x, L.y, L.z, __uint_as_float(sample));
            }
        }
    }
}

// totals: stats[0] += paths, stats[1] += rays, stats[3] += 1 batch
__global__ void k_tally(const uint32_t* q, uint32_t ray_depth, unsigned long long* stats) {
    unsigned long long rays = 0;
    for (uint32_t b = 0; b < ray_depth; ++b) rays += q[b];
    stats[0] += q[0];
    stats[1] += rays;
    stats[3] += 1;
}

// Scene::Render's per-pixel tail (src/scene.cpp:201, 227-228, 247): mean, AcesTonemap,
// GammaCorrected, Color::toUInts (src/color.cpp:26-49)
RT_D float aces(float x) {
    const float a = 2.51f, b = 0.03f, c = 2.43f, d = 0.59f, e = 0.14f;
    float num = __fmul_rn(x, __fadd_rn(__fmul_rn(a, x), b));
    float den = __fadd_rn(__fmul_rn(x, __fadd_rn(__fmul_rn(c, x), d)), e);
    float y = __fdiv_rn(num, den);
    y = (y < 1.f) ? y : 1.f;   // std::min(1.f, y): NaN -> 1
    y = (y < 0.f) ? 0.f : y;   // std::max(y, 0.f)
    return y;
}
RT_D unsigned char to_u8(float y) {
    float g = powf(y, 0.454545468f);  // (float)(1. / 2.2)
    return (unsigned char)(int)roundf(__fmul_rn(255.f, g));
}
__global__ void __launch_bounds__(256) k_resolve(const float* accum, float inv_samples, uint32_t nvalues, uint8_t* out) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nvalues; i += gridDim.x * blockDim.x)
        out[i] = to_u8(aces(__fmul_rn(inv_samples, accum[i])));
}
__global__ void __launch_bounds__(256) k_tonemap(const float* rgb, uint32_t nvalues, uint8_t* out) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nvalues; i += gridDim.x * blockDim.x)
        out[i] = to_u8(aces(rgb[i]));
}

// ------------------------------------------------------------------------------- probes
__global__ void __launch_bounds__(128) k_intersect_batch(DevScene S, long n, const float* o, const float* d, int mode,
                                                          int32_t* id, float* t, float* nrm, int32_t* interior,
                                                          unsigned long long* stats) {
    uint32_t visits = 0, fallbacks = 0;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        vec3 ro = mk3(o[3 * i], o[3 * i + 1], o[3 * i + 2]), rd = mk3(d[3 * i], d[3 * i + 1], d[3 * i + 2]);
        SceneHit h = (mode == 1) ? scene_intersect<1>(S, ro, rd, &visits, nullptr, &fallbacks) : scene_intersect<0>(S, ro, rd, &visits, nullptr, &fallbacks);
        id[i] = h.id;
        t[i] = h.t;
        nrm[3 * i] = h.n.x; nrm[3 * i + 1] = h.n.y; nrm[3 * i + 2] = h.n.z;
        interior[i] = h.interior;
    }
    if (stats) {
        if (visits) atomicAdd(stats + 4, (unsigned long long)visits);
        if (fallbacks) atomicAdd(stats + 5, (unsigned long long)fallbacks);
    }
}
__global__ void k_primitive_batch(DevScene S, uint32_t prim, long n, const float* o, const float* d, int32_t* hit, float* t,
                                  float* nrm, int32_t* interior) {
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        vec3 ro = mk3(o[3 * i], o[3 * i + 1], o[3 * i + 2]), rd = mk3(d[3 * i], d[3 * i + 1], d[3 * i + 2]);
        Isect is;
        bool ok = prim_intersect(S, prim, ro, rd, is);
        hit[i] = ok;
        t[i] = ok ? is.t : 0.f;
        nrm[3 * i] = ok ? is.n.x : 0.f; nrm[3 * i + 1] = ok ? is.n.y : 0.f; nrm[3 * i + 2] = ok ? is.n.z : 0.f;
        interior[i] = ok ? is.interior : 0;
    }
}
__global__ void k_camera_batch(DevScene S, long n, const float* xy, float* o, float* d) {
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        vec3 ro, rd;
        camera_ray(S, xy[2 * i], xy[2 * i + 1], ro, rd);
        o[3 * i] = ro.x; o[3 * i + 1] = ro.y; o[3 * i + 2] = ro.z;
        d[3 * i] = rd.x; d[3 * i + 1] = rd.y; d[3 * i + 2] = rd.z;
    }
}
__global__ void k_pdf_batch(DevScene S, long n, const float* x, const float* nr, const float* d, float* pdf) {
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
        pdf[i] = mix_pdf(S, mk3(x[3 * i], x[3 * i + 1], x[3 * i + 2]), mk3(nr[3 * i], nr[3 * i + 1], nr[3 * i + 2]),
                         mk3(d[3 * i], d[3 * i + 1], d[3 * i + 2]));
}
__global__ void k_sample_batch(DevScene S, long n, const float* x, const float* nr, uint32_t seed, uint32_t sample,
                               uint32_t bounce, float* dir) {
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        Rng g{seed, (uint32_t)i, sample, bounce};
        vec3 r = mix_sample(S, g, mk3(x[3 * i], x[3 * i + 1], x[3 * i + 2]), mk3(nr[3 * i], nr[3 * i + 1], nr[3 * i + 2]));
        dir[3 * i] = r.x; dir[3 * i + 1] = r.y; dir[3 * i + 2] = r.z;
    }
}

// ------------------------------------------------------------------------------- launchers
static int grid_for(uint64_t n, int block, int sms, int per_sm) {
    uint64_t want = (n + block - 1) / block;
    uint64_t cap = (uint64_t)sms * per_sm;
    if (want < 1) want = 1;
    return (int)(want < cap ? want : cap);
}

void launch_generate(const LaunchCtx& c, const DevScene& S, PathSoA P, uint32_t* q, uint64_t first_path, uint32_t count,
                     uint32_t seed, uint32_t sample_begin) {
    k_generate<<<grid_for(count, 256, c.sms, 8), 256, 0, c.stream>>>(S, P, q, first_path, count, seed, sample_begin);
}
void launch_extend(const LaunchCtx& c, const DevScene& S, PathSoA P, HitSoA H, const uint32_t* qcount, uint32_t max_count,
                   int mode, bool count_visits, unsigned long long* stats) {
    int grid = grid_for(max_count, 128, c.sms, 16);
    if (mode == 1) k_extend<1, false><<<grid, 128, 0, c.stream>>>(S, P, H, qcount, stats);
    else if (count_visits) k_extend<0, true><<<grid, 128, 0, c.stream>>>(S, P, H, qcount, stats);
    else k_extend<0, false><<<grid, 128, 0, c.stream>>>(S, P, H, qcount, stats);
}
void launch_shade(const LaunchCtx& c, const DevScene& S, PathSoA P, HitSoA H, PathSoA N, const uint32_t* qin, uint32_t* qout,
                  uint32_t max_count, float* accum, uint32_t bounce, uint32_t seed) {
    k_shade<<<grid_for(max_count, 256, c.sms, 8), 256, 0, c.stream>>>(S, P, H, N, qin, qout, accum, bounce, seed);
}
void launch_tally(const LaunchCtx& c, const uint32_t* q, uint32_t ray_depth, unsigned long long* stats) {
    k_tally<<<1, 1, 0, c.stream>>>(q, ray_depth, stats);
}
void launch_resolve(const LaunchCtx& c, const float* accum, float inv_samples, uint32_t nvalues, uint8_t* out) {
    k_resolve<<<grid_for(nvalues, 256, c.sms, 8), 256, 0, c.stream>>>(accum, inv_samples, nvalues, out);
}
void launch_tonemap(const LaunchCtx& c, const float* rgb, uint32_t nvalues, uint8_t* out) {
    k_tonemap<<<grid_for(nvalues, 256, c.sms, 8), 256, 0, c.stream>>>(rgb, nvalues, out);
}
void launch_intersect_batch(const LaunchCtx& c, const DevScene& S, long n, const float* o, const float* d, int mode, int32_t* id,
                            float* t, float* nrm, int32_t* interior, unsigned long long* stats) {
    k_intersect_batch<<<grid_for((uint64_t)n, 128, c.sms, 16), 128, 0, c.stream>>>(S, n, o, d, mode, id, t, nrm, interior, stats);
}
void launch_primitive_batch(const LaunchCtx& c, const DevScene& S, uint32_t prim, long n, const float* o, const float* d,
                            int32_t* hit, float* t, float* nrm, int32_t* interior) {
    k_primitive_batch<<<grid_for((uint64_t)n, 128, c.sms, 16), 128, 0, c.stream>>>(S, prim, n, o, d, hit, t, nrm, interior);
}
void launch_camera_batch(const LaunchCtx& c, const DevScene& S, long n, const float* xy, float* o, float* d) {
    k_camera_batch<<<grid_for((uint64_t)n, 128, c.sms, 16), 128, 0, c.stream>>>(S, n, xy, o, d);
}
void launch_pdf_batch(const LaunchCtx& c, const DevScene& S, long n, const float* x, const float* nr, const float* d, float* pdf) {
    k_pdf_batch<<<grid_for((uint64_t)n, 128, c.sms, 16), 128, 0, c.stream>>>(S, n, x, nr, d, pdf);
}
void launch_sample_batch(const LaunchCtx& c, const DevScene& S, long n, const float* x, const float* nr, uint32_t seed,
                         uint32_t sample, uint32_t bounce, float* dir) {
    k_sample_batch<<<grid_for((uint64_t)n, 128, c.sms, 16), 128, 0, c.stream>>>(S, n, x, nr, seed, sample, bounce, dir);
}

}  // namespace rtc
