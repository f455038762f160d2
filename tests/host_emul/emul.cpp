// emul.cpp -- TEST-ONLY host compilation of csrc/rt_device.cuh.
//
// There is no GPU in the development container, so the traversal / replay / shading LOGIC of
// the device code is exercised here by compiling the very same header with g++ against a
// handful of intrinsic shims, and comparing it with the oracle (tests/test_host_emulation.py).
// This file is never part of librtc_b200.so and nothing in the product calls it: the shipped
// library has no CPU path.
#define _GNU_SOURCE 1
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <fstream>
#include <sstream>
#include <string>

template <class T> static inline T __ldg(const T* p) { return *p; }
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
static inline uint32_t __float_as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
static inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
static inline int __clz(uint32_t x) { return x ? __builtin_clz(x) : 32; }
static inline float __fdividef(float a, float b) { return a / b; }
static inline float rsqrtf(float a) { return 1.0f / sqrtf(a); }
#define __logf logf        /* glibc declares __logf / __sincosf itself: map the CUDA fast intrinsics by macro */
#define __sincosf sincosf

#include "rt_device.cuh"

using namespace rtc;

struct EmuScene {
    HostScene host;
    DevScene dev;
};

extern "C" {

void* emu_scene_load(const char* path) {
    std::ifstream in(path, std::ios::binary);
    if (!in) return nullptr;
    std::ostringstream ss;
    ss << in.rdbuf();
    EmuScene* e = new EmuScene();
    e->host.parse(ss.str());
    e->host.init();
    const FlatScene& F = e->host.flat;
    DevScene& S = e->dev;
    memset(&S, 0, sizeof S);
    S.geo0 = (const float4*)F.geo0.data(); S.geo1 = (const float4*)F.geo1.data(); S.geo2 = (const float4*)F.geo2.data();
    S.xf_pos = (const float4*)F.xf_pos.data(); S.xf_rot = (const float4*)F.xf_rot.data();
    S.mat0 = (const float4*)F.mat0.data(); S.mat1 = (const float4*)F.mat1.data();
    S.inodes = (const float4*)F.inodes.data(); S.rnodes = (const float4*)F.rnodes.data();
    S.rmeta = (const uint4*)F.rmeta.data(); S.lca = F.lca.data(); S.lights = F.lights.data();
    S.ubox = (const float4*)F.ubox.data();
    S.planes = (const float4*)F.planes.data();
    S.plights = (const float4*)F.plights.data();
    fill_dev_scalars(e->host, S);
    return e;
}
void emu_scene_free(void* h) { delete (EmuScene*)h; }

void emu_intersect(void* h, long n, const float* o, const float* d, int mode, int32_t* id, float* t, float* nrm,
                   int32_t* interior, uint64_t* stats /* visits, fallbacks */) {
    EmuScene* e = (EmuScene*)h;
    uint64_t visits = 0, fallbacks = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : visits, fallbacks)
    for (long i = 0; i < n; ++i) {
        uint32_t v = 0, f = 0;
        vec3 ro = mk3(o[3 * i], o[3 * i + 1], o[3 * i + 2]), rd = mk3(d[3 * i], d[3 * i + 1], d[3 * i + 2]);
        SceneHit hit = mode == 1 ? scene_intersect<1>(e->dev, ro, rd, &v, nullptr, &f) : scene_intersect<0>(e->dev, ro, rd, &v, nullptr, &f);
        id[i] = hit.id; t[i] = hit.t;
        nrm[3 * i] = hit.n.x; nrm[3 * i + 1] = hit.n.y; nrm[3 * i + 2] = hit.n.z;
        interior[i] = hit.interior;
        visits += v; fallbacks += f;
    }
    if (stats) { stats[0] = visits; stats[1] = fallbacks; }
}
void emu_camera_rays(void* h, long n, const float* xy, float* o, float* d) {
    EmuScene* e = (EmuScene*)h;
    for (long i = 0; i < n; ++i) {
        vec3 ro, rd;
        camera_ray(e->dev, xy[2 * i], xy[2 * i + 1], ro, rd);
        o[3 * i] = ro.x; o[3 * i + 1] = ro.y; o[3 * i + 2] = ro.z;
        d[3 * i] = rd.x; d[3 * i + 1] = rd.y; d[3 * i + 2] = rd.z;
    }
}
void emu_mix_pdf(void* h, long n, const float* x, const float* nr, const float* d, float* pdf) {
    EmuScene* e = (EmuScene*)h;
    for (long i = 0; i < n; ++i)
        pdf[i] = mix_pdf(e->dev, mk3(x[3 * i], x[3 * i + 1], x[3 * i + 2]), mk3(nr[3 * i], nr[3 * i + 1], nr[3 * i + 2]),
                         mk3(d[3 * i], d[3 * i + 1], d[3 * i + 2]));
}
void emu_mix_sample(void* h, long n, const float* x, const float* nr, uint32_t seed, uint32_t sample, uint32_t bounce, float* dir) {
    EmuScene* e = (EmuScene*)h;
    for (long i = 0; i < n; ++i) {
        Rng g{seed, (uint32_t)i, sample, bounce};
        vec3 r = mix_sample(e->dev, g, mk3(x[3 * i], x[3 * i + 1], x[3 * i + 2]), mk3(nr[3 * i], nr[3 * i + 1], nr[3 * i + 2]));
        dir[3 * i] = r.x; dir[3 * i + 1] = r.y; dir[3 * i + 2] = r.z;
    }
}
}
