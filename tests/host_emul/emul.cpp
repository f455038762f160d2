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
#include <vector>
#include <algorithm>

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

#include "course_device.cuh"

using namespace rtc;

struct EmuScene {
    HostScene host;
    DevScene dev;
};

extern "C" {

static void* emu_build(const std::string& text, int dialect) {
    EmuScene* e = new EmuScene();
    e->host.dialect = dialect;
    try {
        e->host.parse(text);
        e->host.init();
    } catch (const std::exception&) {
        delete e;
        return nullptr;
    }
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
void* emu_scene_load(const char* path) {
    std::ifstream in(path, std::ios::binary);
    if (!in) return nullptr;
    std::ostringstream ss;
    ss << in.rdbuf();
    return emu_build(ss.str(), DIALECT_HW5);
}
void* emu_scene_parse_dialect(const char* text, long len, int dialect) { return emu_build(std::string(text, (size_t)len), dialect); }
// the deterministic dialects (course_device.cuh): the whole frame as linear colours, (H, W, 3) floats
int emu_frame_linear(void* h, float* out) {
    EmuScene* e = (EmuScene*)h;
    const DevScene& S = e->dev;
    if (S.dialect != DIALECT_HW1 && S.dialect != DIALECT_HW2) return -1;
    const long npix = (long)S.width * S.height;
#pragma omp parallel for schedule(dynamic, 64)
    for (long i = 0; i < npix; ++i) {
        vec3 c = S.dialect == DIALECT_HW1 ? raycast_pixel_hw1(S, (uint32_t)i) : whitted_pixel_hw2(S, (uint32_t)i);
        out[3 * i] = c.x; out[3 * i + 1] = c.y; out[3 * i + 2] = c.z;
    }
    return 0;
}
void emu_scene_free(void* h) { delete (EmuScene*)h; }

void emu_intersect(void* h, long n, const float* o, const float* d, int mode, int32_t* id, float* t, float* nrm,
                   int32_t* interior, uint64_t* stats /* visits, fallbacks, primitive tests */) {
    EmuScene* e = (EmuScene*)h;
    uint64_t visits = 0, fallbacks = 0, tests = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : visits, fallbacks, tests)
    for (long i = 0; i < n; ++i) {
        uint32_t v = 0, f = 0, pt = 0;
        vec3 ro = mk3(o[3 * i], o[3 * i + 1], o[3 * i + 2]), rd = mk3(d[3 * i], d[3 * i + 1], d[3 * i + 2]);
        SceneHit hit = mode == 1 ? scene_intersect<1>(e->dev, ro, rd, &v, &pt, &f) : scene_intersect<0>(e->dev, ro, rd, &v, &pt, &f);
        id[i] = hit.id; t[i] = hit.t;
        nrm[3 * i] = hit.n.x; nrm[3 * i + 1] = hit.n.y; nrm[3 * i + 2] = hit.n.z;
        interior[i] = hit.interior;
        visits += v; fallbacks += f; tests += pt;
    }
    if (stats) { stats[0] = visits; stats[1] = fallbacks; stats[2] = tests; }
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

// ---- the feature-specialised instantiations k_shade uses (FEAT = 0, FE_SPECULAR, FE_ALL): same results on a
//      scene whose features they cover
uint32_t emu_scene_features(void* h) { return ((EmuScene*)h)->dev.features; }
}  // extern "C"
template <uint32_t FEAT>
static void shade_parts(const DevScene& S, long n, const float* x, const float* nr, uint32_t seed, float* dir, float* pdf,
                        const int32_t* prim, const float* o, const float* d, float* tn) {
    for (long i = 0; i < n; ++i) {
        Rng g{seed, (uint32_t)i, 3u, 1u};
        vec3 X = mk3(x[3 * i], x[3 * i + 1], x[3 * i + 2]), Nn = mk3(nr[3 * i], nr[3 * i + 1], nr[3 * i + 2]);
        vec3 r = mix_sample<FEAT>(S, g, X, Nn);
        dir[3 * i] = r.x; dir[3 * i + 1] = r.y; dir[3 * i + 2] = r.z;
        pdf[i] = mix_pdf<FEAT>(S, X, Nn, r);
        Isect is;
        is.t = -1.f; is.n = mk3(0, 0, 0); is.interior = 0;
        bool ok = prim[i] >= 0 && prim_intersect<false, FEAT>(S, (uint32_t)prim[i], mk3(o[3 * i], o[3 * i + 1], o[3 * i + 2]),
                                                               mk3(d[3 * i], d[3 * i + 1], d[3 * i + 2]), is);
        tn[4 * i] = ok ? is.t : -1.f; tn[4 * i + 1] = is.n.x; tn[4 * i + 2] = is.n.y; tn[4 * i + 3] = is.n.z;
        float cd; int id;
        closest_plane<FEAT>(S, X, r, cd, id);
        tn[4 * i] += 0.f * cd;  // keep the call alive without changing the output layout
        pdf[i] = id >= 0 ? pdf[i] : -pdf[i];
    }
}
extern "C" {
int emu_shade_parts(void* h, uint32_t feat, long n, const float* x, const float* nr, uint32_t seed, float* dir, float* pdf,
                    const int32_t* prim, const float* o, const float* d, float* tn) {
    const DevScene& S = ((EmuScene*)h)->dev;
    if (feat == 0) shade_parts<0>(S, n, x, nr, seed, dir, pdf, prim, o, d, tn);
    else if (feat == FE_SPECULAR) shade_parts<FE_SPECULAR>(S, n, x, nr, seed, dir, pdf, prim, o, d, tn);
    else if (feat == FE_ALL) shade_parts<FE_ALL>(S, n, x, nr, seed, dir, pdf, prim, o, d, tn);
    else return -1;
    return 0;
}

// ---- development probe: where the index traversal spends its node visits.  Per node HEIGHT (0 = all children are
//      leaves): visits, children whose box test passed, children that also passed the cone test, per call.
int emu_index_profile(void* h, long n, const float* o, const float* d, uint64_t* out /* [32][3] */) {
    EmuScene* e = (EmuScene*)h;
    const DevScene& S = e->dev;
    const FlatScene& F = e->host.flat;
    const size_t nn = F.inodes.size() / kIndexNodeF4;
    std::vector<int> height(nn, -1);
    // children are emitted after their parent: a reverse sweep sees them first
    for (size_t i = nn; i-- > 0;) {
        int hh = 0;
        for (int c = 0; c < (int)kNodeWidth; ++c) {
            const uint32_t* w = (const uint32_t*)&F.inodes[kIndexNodeF4 * i + kIndexBlockF4 * (c / 4)];
            uint32_t r = w[12 + (c & 3)];
            if (r == IREF_NONE || (r & IREF_LEAF)) continue;
            hh = std::max(hh, height[r & IREF_NODE_MASK] + 1);
        }
        height[i] = hh;
    }
    for (int i = 0; i < 32 * 3; ++i) out[i] = 0;
    if (S.iroot == IREF_NONE || (S.iroot & IREF_LEAF)) return 0;
    for (long i = 0; i < n; ++i) {
        vec3 ro = mk3(o[3 * i], o[3 * i + 1], o[3 * i + 2]), rd = mk3(d[3 * i], d[3 * i + 1], d[3 * i + 2]);
        vec3 inv = ray_inv(rd), oi = ro * inv;
        ConeDir dn = cone_dir(rd), none = dn;
        none.xy = 0x7E007E00u; none.zz = 0x7E007E00u;  // NaN halves: nothing is culled by the cones
        std::vector<uint32_t> st{S.iroot};
        while (!st.empty()) {
            uint32_t node = st.back();
            st.pop_back();
            NodeVisit box = index_visit(S, node, inv, oi, none), both = index_visit(S, node, inv, oi, dn);
            int hh = std::min(height[node & IREF_NODE_MASK], 31);
            out[3 * hh] += 1;
            for (int c = 0; c < (int)kNodeWidth; ++c) {
                out[3 * hh + 1] += box.hit[c];
                out[3 * hh + 2] += both.hit[c];
                if (both.hit[c] && !(both.ref[c] & IREF_LEAF)) st.push_back(both.ref[c]);
            }
        }
    }
    return (int)nn;
}

// ---- structure of the index tree: out = {nodes, leaf slots, leaf slots seen twice, units missing, nodes with fewer
//      than 2 used slots (other than the root), child boxes (fp16, as the device decodes them) that fail to contain
//      the exact box of a reference leaf below them, deepest level}
int emu_index_check(void* h, uint64_t* out /* [7] */) {
    EmuScene* e = (EmuScene*)h;
    const DevScene& S = e->dev;
    const FlatScene& F = e->host.flat;
    for (int i = 0; i < 7; ++i) out[i] = 0;
    const size_t nn = F.inodes.size() / kIndexNodeF4;
    out[0] = nn;
    if (S.iroot == IREF_NONE || (S.iroot & IREF_LEAF)) return 0;
    std::vector<uint8_t> seen(e->host.prims.size(), 0);
    struct Item { uint32_t node; int depth; float mn[3], mx[3]; bool bounded; };
    std::vector<Item> st;
    st.push_back(Item{S.iroot, 1, {0, 0, 0}, {0, 0, 0}, false});
    while (!st.empty()) {
        Item it = st.back();
        st.pop_back();
        out[6] = std::max<uint64_t>(out[6], (uint64_t)it.depth);
        int used = 0;
        for (uint32_t blk = 0; blk < kNodeWidth / 4; ++blk) {
        const float4* nd = S.inodes + kIndexNodeF4 * (size_t)(it.node & IREF_NODE_MASK) + kIndexBlockF4 * blk;
        const float4 q0 = nd[0], q1 = nd[1], q2 = nd[2], q3 = nd[3];
        float2 cx[2] = {unpack_half2(q0.x), unpack_half2(q0.y)}, cy[2] = {unpack_half2(q0.z), unpack_half2(q0.w)};
        float2 cz[2] = {unpack_half2(q1.x), unpack_half2(q1.y)}, hx[2] = {unpack_half2(q1.z), unpack_half2(q1.w)};
        float2 hy[2] = {unpack_half2(q2.x), unpack_half2(q2.y)}, hz[2] = {unpack_half2(q2.z), unpack_half2(q2.w)};
        uint32_t refs[4] = {__float_as_uint(q3.x), __float_as_uint(q3.y), __float_as_uint(q3.z), __float_as_uint(q3.w)};
        for (int c = 0; c < 4; ++c) {
            if (refs[c] == IREF_NONE) continue;
            ++used;
            auto pick = [&](float2* a) { return (c & 1) ? a[c >> 1].y : a[c >> 1].x; };
            float C[3] = {pick(cx), pick(cy), pick(cz)}, H[3] = {pick(hx), pick(hy), pick(hz)};
            Item ch{0, it.depth + 1, {C[0] - H[0], C[1] - H[1], C[2] - H[2]}, {C[0] + H[0], C[1] + H[1], C[2] + H[2]}, true};
            if (it.bounded)  // a child also lies inside everything above it
                for (int a = 0; a < 3; ++a) { ch.mn[a] = std::max(ch.mn[a], it.mn[a]); ch.mx[a] = std::min(ch.mx[a], it.mx[a]); }
            if (refs[c] & IREF_LEAF) {
                ++out[1];
                uint32_t first = refs[c] & 0xFFFFFFu;
                if (first >= seen.size()) { ++out[3]; continue; }
                if (seen[first]) ++out[2];
                seen[first] = 1;
                const f4 bmn = F.ubox[2 * (size_t)first], bmx = F.ubox[2 * (size_t)first + 1];
                const float emn[3] = {bmn.x, bmn.y, bmn.z}, emx[3] = {bmx.x, bmx.y, bmx.z};
                for (int a = 0; a < 3; ++a)
                    if (!(ch.mn[a] <= emn[a] && emx[a] <= ch.mx[a])) { ++out[5]; break; }
            } else {
                ch.node = refs[c];
                st.push_back(ch);
            }
        }
        }
        if (used < 2 && it.node != S.iroot) ++out[4];
    }
    // every reference leaf (unit) must have been reached: units are the leaves of the reference tree
    uint64_t units = 0;
    for (size_t v = 0; v < e->host.nodes.size(); ++v)
        if (e->host.nodes[v].left == UINT32_MAX && e->host.nodes[v].count > 0) {
            ++units;
            if (!seen[e->host.nodes[v].first]) ++out[3];
        }
    return (int)units;
}

// FNV-1a over the bytes of the flattened index tree (and its root reference): lets a test or a developer confirm
// that a change to the host builder left the device data byte-identical
uint64_t emu_index_hash(void* h) {
    EmuScene* e = (EmuScene*)h;
    const FlatScene& F = e->host.flat;
    uint64_t x = 1469598103934665603ull;
    auto eat = [&](const void* p, size_t n) { const unsigned char* b = (const unsigned char*)p; for (size_t i = 0; i < n; ++i) { x ^= b[i]; x *= 1099511628211ull; } };
    eat(F.inodes.data(), F.inodes.size() * sizeof(f4));
    eat(&F.iroot, sizeof F.iroot);
    return x;
}
}
