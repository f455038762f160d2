// refprobe.cpp -- thin C-ABI probe over the UNMODIFIED reference (TEST INFRASTRUCTURE ONLY).
//
// Linked by oracle/Makefile against the reference's own sources, compiled where they lie
// under /root/reference/hw5 (nothing of the reference is copied into this repo).  It exposes
// the reference's hot-path functions (Scene::RayIntersection, Primitive::Intersect,
// Camera::GetToRay, Distribution::Pdf, AcesTonemap/GammaCorrected/toUInts, Scene::Sample)
// batch-wise, so that tools/make_golden.py can record golden vectors from it and
// tests/test_oracle_golden.py can pin oracle/rt_oracle.c against the real thing.
// Scene's private members are reached with the usual "#define private public" test trick,
// applied only after every standard/glm header has already been included.
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <functional>
#include <iostream>
#include <memory>
#include <optional>
#include <random>
#include <sstream>
#include <string>
#include <thread>
#include <variant>
#include <vector>
#include <omp.h>

#include "distributions.h"   // pulls primitives.h, color.h, quaternion.h, point.h, glm
#define private public
#include "bvh.h"
#include "scene.h"
#undef private

extern "C" {

void* ref_scene_load(const char* path) {
    std::ifstream in(path);
    if (!in) return nullptr;
    Scene* s = new Scene();
    s->cam.pos = {0, 0, 0};
    s->Load(in);
    s->InitScene();
    return s;
}
void ref_scene_free(void* h) { delete static_cast<Scene*>(h); }

void ref_scene_info(void* h, uint32_t out[8]) {
    Scene* s = static_cast<Scene*>(h);
    uint32_t nb = 0;
    for (auto& p : s->primitives) nb += p.primitive_type != PRIMITIVE_TYPE::PLANE;
    out[0] = s->cam.width; out[1] = s->cam.height; out[2] = s->ray_depth; out[3] = s->samples;
    out[4] = (uint32_t)s->primitives.size(); out[5] = nb; out[6] = (uint32_t)s->scene_bvh.nodes.size();
    out[7] = (uint32_t)std::get<2>(s->mix_distrib.data).size();
}
void ref_scene_override(void* h, int width, int height, int samples, int ray_depth) {
    Scene* s = static_cast<Scene*>(h);
    if (width >= 0) s->cam.width = width;
    if (height >= 0) s->cam.height = height;
    if (samples >= 0) s->samples = samples;
    if (ray_depth >= 0) s->ray_depth = ray_depth;
}
void ref_scene_prims(void* h, int32_t* tm, float* d) {
    Scene* s = static_cast<Scene*>(h);
    for (size_t i = 0; i < s->primitives.size(); ++i) {
        const Primitive& p = s->primitives[i];
        tm[2 * i] = (int)p.primitive_type; tm[2 * i + 1] = (int)p.material;
        float* o = d + 26 * i;
        o[0] = p.col.rgb.x; o[1] = p.col.rgb.y; o[2] = p.col.rgb.z;
        o[3] = p.emission.rgb.x; o[4] = p.emission.rgb.y; o[5] = p.emission.rgb.z;
        o[6] = p.pos.x; o[7] = p.pos.y; o[8] = p.pos.z;
        o[9] = p.rotator.x; o[10] = p.rotator.y; o[11] = p.rotator.z; o[12] = p.rotator.w;
        o[13] = p.ior;
        o[14] = p.dop_data.x; o[15] = p.dop_data.y; o[16] = p.dop_data.z;
        bool tri = p.primitive_type == PRIMITIVE_TYPE::TRIANGLE;
        o[17] = tri ? p.dop_data1.x : 0; o[18] = tri ? p.dop_data1.y : 0; o[19] = tri ? p.dop_data1.z : 0;
        o[20] = tri ? p.dop_data2.x : 0; o[21] = tri ? p.dop_data2.y : 0; o[22] = tri ? p.dop_data2.z : 0;
        o[23] = o[24] = o[25] = 0;
    }
}
void ref_scene_nodes(void* h, float* aabb, uint32_t* links) {
    Scene* s = static_cast<Scene*>(h);
    auto& nodes = s->scene_bvh.nodes;
    for (size_t i = 0; i < nodes.size(); ++i) {
        const NODE_t& n = nodes[i];
        aabb[6 * i] = n.aabb.aabb_min.x; aabb[6 * i + 1] = n.aabb.aabb_min.y; aabb[6 * i + 2] = n.aabb.aabb_min.z;
        aabb[6 * i + 3] = n.aabb.aabb_max.x; aabb[6 * i + 4] = n.aabb.aabb_max.y; aabb[6 * i + 5] = n.aabb.aabb_max.z;
        links[4 * i] = n.left_child; links[4 * i + 1] = n.right_child;
        links[4 * i + 2] = n.first_primitive_id; links[4 * i + 3] = n.primitive_count;
    }
}
uint32_t ref_scene_root(void* h) { return static_cast<Scene*>(h)->scene_bvh.root_; }

void ref_intersect(void* h, long n, const float* o, const float* d, int32_t* id, float* t, float* normal, int32_t* interior) {
    Scene* s = static_cast<Scene*>(h);
#pragma omp parallel for schedule(dynamic, 256)
    for (long i = 0; i < n; ++i) {
        Ray r({o[3 * i], o[3 * i + 1], o[3 * i + 2]}, {d[3 * i], d[3 * i + 1], d[3 * i + 2]});
        ray_intersection_t hit = s->RayIntersection(r);
        bool ok = hit.id != -1;
        id[i] = hit.id;
        t[i] = ok ? hit.isec.t : 0.f;
        normal[3 * i] = ok ? hit.isec.normal.x : 0.f;
        normal[3 * i + 1] = ok ? hit.isec.normal.y : 0.f;
        normal[3 * i + 2] = ok ? hit.isec.normal.z : 0.f;
        interior[i] = ok ? (int)hit.isec.interior : 0;
    }
}
void ref_primitive_intersect(void* h, int prim, long n, const float* o, const float* d,
                             int32_t* hit, float* t, float* normal, int32_t* interior) {
    Scene* s = static_cast<Scene*>(h);
    for (long i = 0; i < n; ++i) {
        Ray r({o[3 * i], o[3 * i + 1], o[3 * i + 2]}, {d[3 * i], d[3 * i + 1], d[3 * i + 2]});
        auto is = s->primitives[prim].Intersect(r);
        hit[i] = is.has_value();
        t[i] = is ? is->t : 0.f;
        normal[3 * i] = is ? is->normal.x : 0.f; normal[3 * i + 1] = is ? is->normal.y : 0.f; normal[3 * i + 2] = is ? is->normal.z : 0.f;
        interior[i] = is ? (int)is->interior : 0;
    }
}
void ref_camera_rays(void* h, long n, const float* xy, float* o, float* d) {
    Scene* s = static_cast<Scene*>(h);
    for (long i = 0; i < n; ++i) {
        Ray r = s->cam.GetToRay(xy[2 * i], xy[2 * i + 1]);
        o[3 * i] = r.o.x; o[3 * i + 1] = r.o.y; o[3 * i + 2] = r.o.z;
        d[3 * i] = r.d.x; d[3 * i + 1] = r.d.y; d[3 * i + 2] = r.d.z;
    }
}
void ref_mix_pdf(void* h, long n, const float* x, const float* nrm, const float* d, float* pdf) {
    Scene* s = static_cast<Scene*>(h);
    for (long i = 0; i < n; ++i)
        pdf[i] = s->mix_distrib.Pdf({x[3 * i], x[3 * i + 1], x[3 * i + 2]}, {nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]},
                                    {d[3 * i], d[3 * i + 1], d[3 * i + 2]});
}
// Distribution::Sample with the reference's own minstd stream seeded by (seed0 + i): used for
// STATISTICAL comparison of the sampled direction distribution only.
void ref_mix_sample(void* h, long n, const float* x, const float* nrm, uint32_t seed0, float* dir) {
    Scene* s = static_cast<Scene*>(h);
    for (long i = 0; i < n; ++i) {
        std::minstd_rand rnd(seed0 + (uint32_t)i);
        std::uniform_real_distribution<float> uniform01{0.f, 1.f};
        std::normal_distribution<float> normal01{0.f, 1.f};
        RANDOM_t random{rnd, uniform01, normal01};
        glm::vec3 r = s->mix_distrib.Sample(random, {x[3 * i], x[3 * i + 1], x[3 * i + 2]}, {nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]});
        dir[3 * i] = r.x; dir[3 * i + 1] = r.y; dir[3 * i + 2] = r.z;
    }
}
void ref_tonemap_u8(long npix, const float* rgb, uint8_t* out) {
    for (long i = 0; i < npix; ++i) {
        Color c(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]);
        c = AcesTonemap(c);
        c = GammaCorrected(c);
        unsigned char* u = c.toUInts();
        out[3 * i] = u[0]; out[3 * i + 1] = u[1]; out[3 * i + 2] = u[2];
        delete[] u;
    }
}
// Scene::Render's per-pixel loop (scene.cpp:214-231) stopping BEFORE tonemapping: linear mean
// radiance per pixel, with the reference's own per-pixel minstd_rand(i) streams.
void ref_render_linear(void* h, long pix_begin, long pix_end, float* out_mean, int nthreads) {
    Scene* s = static_cast<Scene*>(h);
    omp_set_num_threads(nthreads > 0 ? nthreads : (int)std::thread::hardware_concurrency());
#pragma omp parallel for schedule(dynamic)
    for (long i = pix_begin; i < pix_end; ++i) {
        std::minstd_rand rnd(i);
        std::uniform_real_distribution<float> uniform01{0.f, 1.f};
        std::normal_distribution<float> normal01{0.f, 1.f};
        RANDOM_t random{rnd, uniform01, normal01};
        unsigned int x = i % s->cam.width, y = i / s->cam.width;
        Color c = s->Sample(random, x, y);
        out_mean[3 * (i - pix_begin)] = c.rgb.x; out_mean[3 * (i - pix_begin) + 1] = c.rgb.y; out_mean[3 * (i - pix_begin) + 2] = c.rgb.z;
    }
}

// The real libstdc++ algorithms the reference's BVH order depends on (bvh.cpp:129,168; scene.cpp:17)
void ref_std_sort_perm(const float* key, int32_t* perm, long first, long last) {
    std::sort(perm + first, perm + last, [key](int32_t a, int32_t b) { return key[a] < key[b]; });
}
long ref_std_partition(int32_t* perm, const uint8_t* pred, long n) {
    return std::partition(perm, perm + n, [pred](int32_t a) { return pred[a] != 0; }) - perm;
}

}  // extern "C"
