/*
 * rt_oracle.c -- CPU oracle (TEST INFRASTRUCTURE ONLY, see rt_oracle.h).
 *
 * Plain-C restatement of /root/reference/hw5 (glm 1.0.0 float math written out by hand).
 * Every function cites the reference file:line it follows.  Compile with
 *   gcc -O2 -ffp-contract=off -fopenmp -fPIC -shared   (see oracle/Makefile)
 * -ffp-contract=off matters: the reference is built for baseline x86-64 (no FMA), and
 * the deterministic functions here are pinned bit-for-bit against it.
 */
#include "rt_oracle.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ vec math (glm) */
typedef struct { float x, y, z; } v3;
typedef struct { float x, y, z, w; } quat;

static inline v3 V(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline v3 vadd(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 vsub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 vmul(v3 a, v3 b) { return V(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline v3 vdiv(v3 a, v3 b) { return V(a.x / b.x, a.y / b.y, a.z / b.z); }
static inline v3 vscale(float k, v3 a) { return V(k * a.x, k * a.y, k * a.z); }
/* glm/detail/func_geometric.inl:48-55 : tmp = a*b; tmp.x + tmp.y + tmp.z */
static inline float vdot(v3 a, v3 b) { v3 t = vmul(a, b); return t.x + t.y + t.z; }
static inline v3 vcross(v3 a, v3 b) {
    return V(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
}
static inline float vlength(v3 a) { return sqrtf(vdot(a, a)); }
/* glm normalize = v * inversesqrt(dot(v,v)), inversesqrt = 1/sqrt (func_geometric.inl:82-90) */
static inline v3 vnormalize(v3 a) { float k = 1.0f / sqrtf(vdot(a, a)); return V(a.x * k, a.y * k, a.z * k); }
static inline float vidx(v3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }
static inline quat qconj(quat q) { quat r = {-q.x, -q.y, -q.z, q.w}; return r; }
/* glm/detail/type_quat.inl:359-366 ; reference quaternion.cpp:5-7 */
static inline v3 qrot(quat q, v3 v) {
    v3 qv = V(q.x, q.y, q.z);
    v3 uv = vcross(qv, v);
    v3 uuv = vcross(qv, uv);
    v3 s = vadd(vscale(q.w, uv), uuv); /* (uv * q.w) + uuv */
    return vadd(v, V(s.x * 2.0f, s.y * 2.0f, s.z * 2.0f));
}
static inline float fminstd(float a, float b) { return (b < a) ? b : a; } /* std::min(a,b) */
static inline float fmaxstd(float a, float b) { return (a < b) ? b : a; } /* std::max(a,b) */

/* ------------------------------------------------------------------ scene types */
enum { PT_PLANE = 1, PT_BOX = 2, PT_ELLIPSOID = 4, PT_TRIANGLE = 8 };   /* primitives.h:13-18 */
enum { MAT_DIFFUSE = 0, MAT_METALLIC = 1, MAT_DIELECTRIC = 2 };         /* materials.h */

typedef struct {
    int type, material, orig;
    v3 col, emission, pos;
    quat rot;
    float ior;
    v3 d0, d1, d2;
} prim_t;

typedef struct { v3 mn, mx; } aabb_t;
typedef struct { aabb_t box; uint32_t left, right, first, count; } node_t;
typedef struct { float t; v3 n; int interior; } isec_t;
typedef struct { isec_t isec; int id; } rayisec_t;
typedef struct { v3 o, d; } ray_t;

/* hw2 light (hw2 src/lights.cpp) */
typedef struct { v3 intensity, pos, att, dir; int directed; } plight_t;

struct orc_scene {
    int dialect;                 /* 1..5: which homework snapshot's Scene::Load / Scene::RayTrace apply */
    v3 ambient;                  /* hw2 AMBIENT_LIGHT */
    plight_t* plights; int nplights, plcap;
    unsigned width, height, ray_depth, samples;
    v3 bg, cam_pos, cam_right, cam_up, cam_forward;
    float fov_x;
    prim_t* prims; int nprims, cap;
    int nbvh;
    node_t* nodes; int nnodes, nodecap; uint32_t root;
    int* lights; int nlights;
};

static const float INF_ = 1e18f;      /* bvh.h:9 */

/* ------------------------------------------------------------------ primitive intersections */
/* primitives.cpp:55-66 */
/* IntersectPlane drops t > 1e5 from hw4 on (hw4 src/primitives.cpp:47, hw5 :58); hw1..hw3 have no limit */
static int isect_plane_tmax(ray_t r, v3 n, float tmax, isec_t* out) {
    float t = -vdot(r.o, n) / vdot(r.d, n);
    if (t > tmax) return 0;
    if (t > 0) {
        if (vdot(r.d, n) >= 0) { out->t = t; out->n = vscale(-1.0f, n); out->interior = 1; return 1; }
        out->t = t; out->n = n; out->interior = 0; return 1;
    }
    return 0;
}
static int isect_plane(ray_t r, v3 n, isec_t* out) { return isect_plane_tmax(r, n, 1e5f, out); }
/* primitives.cpp:70-117 */
static int isect_box(ray_t r, v3 s, isec_t* out) {
    v3 t1v = vdiv(vsub(vscale(-1.f, s), r.o), r.d);
    v3 t2v = vdiv(vsub(s, r.o), r.d);
    float t1x = fminstd(t1v.x, t2v.x), t2x = fmaxstd(t1v.x, t2v.x);
    float t1y = fminstd(t1v.y, t2v.y), t2y = fmaxstd(t1v.y, t2v.y);
    float t1z = fminstd(t1v.z, t2v.z), t2z = fmaxstd(t1v.z, t2v.z);
    float t1 = fmaxstd(fmaxstd(t1x, t1y), t1z);
    float t2 = fminstd(fminstd(t2x, t2y), t2z);
    if (t1 > t2) return 0;
    if (t2 < 0) return 0;
    int interior = t1 < 0;
    float t = interior ? t2 : t1;
    v3 p = vadd(r.o, vscale(t, r.d));
    v3 nrm = vdiv(p, s);
    if (interior) nrm = vscale(-1.0f, nrm);
    float mx = fmaxstd(fmaxstd(fabsf(nrm.x), fabsf(nrm.y)), fabsf(nrm.z));
    if (fabsf(nrm.x) != mx) nrm.x = 0;
    if (fabsf(nrm.y) != mx) nrm.y = 0;
    if (fabsf(nrm.z) != mx) nrm.z = 0;
    out->t = t; out->n = vnormalize(nrm); out->interior = interior;
    return 1;
}
/* primitives.cpp:120-152 */
static int isect_ellipsoid(ray_t r, v3 rad, isec_t* out) {
    v3 dr = vdiv(r.d, rad), orr = vdiv(r.o, rad);
    float a = vdot(dr, dr);
    float b = 2 * vdot(orr, dr);
    float c = vdot(orr, orr) - 1;
    float d = b * b - 4 * a * c;
    if (d <= 0) return 0;
    /* unqualified sqrt(float) in primitives.cpp resolves to ::sqrt(double): the numerator and
       the division are evaluated in double and rounded once (pinned against the reference) */
    float x1 = (float)(((double)(-b) - sqrt((double)d)) / (double)(2 * a));
    float x2 = (float)(((double)(-b) + sqrt((double)d)) / (double)(2 * a));
    if (x1 > x2) { float tmp = x1; x1 = x2; x2 = tmp; }
    if (x2 < 0) return 0;
    int interior = x1 < 0;
    float t = interior ? x2 : x1;
    v3 p = vadd(r.o, vscale(t, r.d));
    v3 nrm = vnormalize(vdiv(p, vmul(rad, rad)));
    if (interior) nrm = vscale(-1.0f, nrm);
    out->t = t; out->n = nrm; out->interior = interior;
    return 1;
}
/* primitives.cpp:155-174.  NOTE (faithful): the plane used is the one through the LOCAL
   ORIGIN with the triangle's normal, not the plane through vertex a. */
static int isect_triangle(ray_t r, v3 a, v3 b, v3 c, isec_t* out) {
    v3 n = vnormalize(vcross(vsub(b, a), vsub(c, a)));
    isec_t is;
    if (!isect_plane(r, n, &is)) return 0;
    v3 p = vadd(r.o, vscale(is.t, r.d));
    if (!(vdot(vcross(vsub(b, a), vsub(p, a)), n) > 0)) return 0;
    if (!(vdot(vcross(vsub(p, a), vsub(c, a)), n) > 0)) return 0;
    if (!(vdot(vcross(vsub(c, b), vsub(p, b)), n) > 0)) return 0;
    *out = is;
    return 1;
}
/* primitives.cpp:14-52 */
static int prim_intersect_ex(const prim_t* pr, ray_t ray, float plane_tmax, isec_t* out) {
    quat qc = qconj(pr->rot);
    ray_t rr;
    rr.o = qrot(qc, vadd(ray.o, vscale(-1.0f, pr->pos)));
    rr.d = qrot(qc, ray.d);
    isec_t is;
    int ok = 0;
    switch (pr->type) {
        case PT_PLANE: ok = isect_plane_tmax(rr, pr->d0, plane_tmax, &is); break;
        case PT_BOX: ok = isect_box(rr, pr->d0, &is); break;
        case PT_ELLIPSOID: ok = isect_ellipsoid(rr, pr->d0, &is); break;
        case PT_TRIANGLE: ok = isect_triangle(rr, pr->d0, pr->d1, pr->d2, &is); break;
        default: return 0;
    }
    if (!ok) return 0;
    out->t = is.t;
    out->n = vnormalize(qrot(pr->rot, is.n));
    out->interior = is.interior;
    return 1;
}
static int prim_intersect(const prim_t* pr, ray_t ray, isec_t* out) { return prim_intersect_ex(pr, ray, 1e5f, out); }

/* ------------------------------------------------------------------ AABB / BVH (bvh.cpp) */
static aabb_t aabb_empty(void) { aabb_t b = {{INF_, INF_, INF_}, {-INF_, -INF_, -INF_}}; return b; }
static void aabb_extend_p(aabb_t* b, v3 p) { /* bvh.cpp:29-34 */
    b->mx.x = fmaxstd(b->mx.x, p.x); b->mn.x = fminstd(b->mn.x, p.x);
    b->mx.y = fmaxstd(b->mx.y, p.y); b->mn.y = fminstd(b->mn.y, p.y);
    b->mx.z = fmaxstd(b->mx.z, p.z); b->mn.z = fminstd(b->mn.z, p.z);
}
static void aabb_extend(aabb_t* b, aabb_t o) { aabb_extend_p(b, o.mx); aabb_extend_p(b, o.mn); } /* :36-39 */
static float aabb_area(aabb_t b) { /* bvh.cpp:23-27 */
    v3 d = vsub(b.mx, b.mn);
    return 2.f * (d.x * d.y + d.x * d.z + d.y * d.z);
}
static float min3f(float a, float b, float c) { /* std::min({a,b,c}) = min_element: first smallest */
    float m = a; if (b < m) m = b; if (c < m) m = c; return m;
}
static float max3f(float a, float b, float c) { /* std::max({..}) = max_element: first largest */
    float m = a; if (m < b) m = b; if (m < c) m = c; return m;
}
/* bvh.cpp:41-87 */
static aabb_t aabb_of_prim(const prim_t* p) {
    v3 mn, mx;
    if (p->type == PT_TRIANGLE) {
        mn = V(min3f(p->d0.x, p->d1.x, p->d2.x), min3f(p->d0.y, p->d1.y, p->d2.y), min3f(p->d0.z, p->d1.z, p->d2.z));
        mx = V(max3f(p->d0.x, p->d1.x, p->d2.x), max3f(p->d0.y, p->d1.y, p->d2.y), max3f(p->d0.z, p->d1.z, p->d2.z));
    } else {
        mn = vscale(-1.f, p->d0); mx = p->d0;
    }
    aabb_t b = aabb_empty();
    for (int mask = 0; mask < 8; ++mask) {
        v3 vtx = V((mask & 1) ? mx.x : mn.x, (mask & 2) ? mx.y : mn.y, (mask & 4) ? mx.z : mn.z);
        aabb_extend_p(&b, qrot(p->rot, vtx));
    }
    b.mn = vadd(b.mn, p->pos);
    b.mx = vadd(b.mx, p->pos);
    return b;
}
/* bvh.cpp:89-93 */
static int aabb_intersect(aabb_t b, ray_t ray, isec_t* out) {
    v3 s = vscale(0.5f, vsub(b.mx, b.mn));
    v3 c = vscale(0.5f, vadd(b.mx, b.mn));
    ray_t r; r.o = vadd(ray.o, vscale(-1.0f, c)); r.d = ray.d;
    return isect_box(r, s, out);
}

/* ---- libstdc++ std::sort (bits/stl_algo.h, introsort + final insertion sort) and
   std::partition (bidirectional version) restated over an index permutation.  The
   reference sorts whole Primitive objects by pos[axis] (bvh.cpp:129-131,168-170); a
   comparison sort's resulting permutation depends only on the comparison outcomes, so
   sorting indices with the same comparator reproduces the reference's final order,
   INCLUDING the order of equal keys (all dragon triangles have pos = 0). */
typedef struct { const float* key; } sortctx;
#define LESS(a, b) (ctx->key[(a)] < ctx->key[(b)])
static void sw(int32_t* a, int32_t* b) { int32_t t = *a; *a = *b; *b = t; }

static void adjust_heap(const sortctx* ctx, int32_t* first, long hole, long len, int32_t value) {
    const long top = hole;
    long child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (LESS(first[child], first[child - 1])) child--;
        first[hole] = first[child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        first[hole] = first[child - 1];
        hole = child - 1;
    }
    long parent = (hole - 1) / 2; /* __push_heap */
    while (hole > top && LESS(first[parent], value)) {
        first[hole] = first[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    first[hole] = value;
}
static void heap_sort_range(const sortctx* ctx, int32_t* first, int32_t* last) {
    long len = last - first; /* __heap_select(first,last,last) == make_heap ; then sort_heap */
    if (len >= 2) {
        long parent = (len - 2) / 2;
        for (;;) {
            int32_t v = first[parent];
            adjust_heap(ctx, first, parent, len, v);
            if (parent == 0) break;
            parent--;
        }
    }
    while (last - first > 1) {
        --last;
        int32_t v = *last; /* __pop_heap(first,last,last) */
        *last = *first;
        adjust_heap(ctx, first, 0, last - first, v);
    }
}
static void move_median_to_first(const sortctx* ctx, int32_t* result, int32_t* a, int32_t* b, int32_t* c) {
    if (LESS(*a, *b)) {
        if (LESS(*b, *c)) sw(result, b);
        else if (LESS(*a, *c)) sw(result, c);
        else sw(result, a);
    } else if (LESS(*a, *c)) sw(result, a);
    else if (LESS(*b, *c)) sw(result, c);
    else sw(result, b);
}
static int32_t* unguarded_partition(const sortctx* ctx, int32_t* first, int32_t* last, int32_t* pivot) {
    for (;;) {
        while (LESS(*first, *pivot)) ++first;
        --last;
        while (LESS(*pivot, *last)) --last;
        if (!(first < last)) return first;
        sw(first, last);
        ++first;
    }
}
static void introsort_loop(const sortctx* ctx, int32_t* first, int32_t* last, long depth_limit) {
    while (last - first > 16) {
        if (depth_limit == 0) { heap_sort_range(ctx, first, last); return; }
        --depth_limit;
        int32_t* mid = first + (last - first) / 2;
        move_median_to_first(ctx, first, first + 1, mid, last - 1);
        int32_t* cut = unguarded_partition(ctx, first + 1, last, first);
        introsort_loop(ctx, cut, last, depth_limit);
        last = cut;
    }
}
static void unguarded_linear_insert(const sortctx* ctx, int32_t* last) {
    int32_t val = *last;
    int32_t* next = last - 1;
    while (LESS(val, *next)) { *last = *next; last = next; --next; }
    *last = val;
}
static void insertion_sort(const sortctx* ctx, int32_t* first, int32_t* last) {
    if (first == last) return;
    for (int32_t* i = first + 1; i != last; ++i) {
        if (LESS(*i, *first)) {
            int32_t val = *i;
            memmove(first + 1, first, (size_t)(i - first) * sizeof(int32_t));
            *first = val;
        } else unguarded_linear_insert(ctx, i);
    }
}
static void std_sort(const sortctx* ctx, int32_t* first, int32_t* last) {
    if (first == last) return;
    long n = last - first, lg = 0;
    while ((n >> (lg + 1)) != 0) lg++; /* std::__lg */
    introsort_loop(ctx, first, last, lg * 2);
    if (last - first > 16) {
        insertion_sort(ctx, first, first + 16);
        for (int32_t* i = first + 16; i != last; ++i) unguarded_linear_insert(ctx, i);
    } else insertion_sort(ctx, first, last);
}
void orc_sort_perm_by_key(const float* key, int32_t* perm, long first, long last) {
    sortctx c = {key};
    std_sort(&c, perm + first, perm + last);
}
/* std::partition, bidirectional iterators (bits/stl_algo.h __partition) */
long orc_partition_flags(int32_t* perm, const uint8_t* pred, long n) {
    int32_t* first = perm; int32_t* last = perm + n;
    for (;;) {
        for (;;) { if (first == last) return first - perm; else if (pred[*first]) ++first; else break; }
        --last;
        for (;;) { if (first == last) return first - perm; else if (!pred[*last]) --last; else break; }
        sw(first, last);
        ++first;
    }
}

typedef struct {
    orc_scene* s;
    int32_t* perm;      /* perm[slot] = index into src prims */
    const prim_t* src;
    aabb_t* pbox;       /* AABB per src prim */
    float* keys[3];     /* pos[axis] per src prim */
    float* cut_qual;
} build_t;

static uint32_t push_node(orc_scene* s, node_t n) {
    if (s->nnodes == s->nodecap) {
        s->nodecap = s->nodecap ? s->nodecap * 2 : 1024;
        s->nodes = (node_t*)realloc(s->nodes, sizeof(node_t) * (size_t)s->nodecap);
    }
    s->nodes[s->nnodes] = n;
    return (uint32_t)s->nnodes++;
}
/* bvh.cpp:105-179 */
static uint32_t init_tree(build_t* B, uint32_t first, uint32_t last) {
    aabb_t box = aabb_empty();
    for (uint32_t i = first; i < last; i++) aabb_extend(&box, B->pbox[B->perm[i]]);
    node_t cur; cur.box = box; cur.first = first; cur.count = last - first;
    cur.left = (uint32_t)-1; cur.right = (uint32_t)-1;
    uint32_t cur_pos = push_node(B->s, cur);
    if (last - first == 1) return cur_pos;

    float optimums[3] = {INF_, INF_, INF_};
    uint32_t cuts[3] = {0, 0, 0};
    float* cq = B->cut_qual;
    for (int axis = 0; axis < 3; ++axis) {
        sortctx c = {B->keys[axis]};
        std_sort(&c, B->perm + first, B->perm + last);
        aabb_t pref = B->pbox[B->perm[first]];
        for (uint32_t cut = first + 1; cut < last; ++cut) {
            cq[cut] = aabb_area(pref) * (float)(cut - first);
            aabb_extend(&pref, B->pbox[B->perm[cut]]);
        }
        aabb_t suf = aabb_empty();
        for (uint32_t cut = last - 1; cut > first; --cut) {
            aabb_extend(&suf, B->pbox[B->perm[cut]]);
            cq[cut] += aabb_area(suf) * (float)(last - cut);
        }
        for (uint32_t cut = first + 1; cut < last; ++cut)
            if (cq[cut] < optimums[axis]) { optimums[axis] = cq[cut]; cuts[axis] = cut; }
    }
    float optimum = min3f(optimums[0], optimums[1], optimums[2]);
    float without_cut = aabb_area(cur.box) * (float)cur.count;
    if (optimum >= without_cut) return cur_pos;
    uint32_t cut = 0;
    for (int axis = 0; axis < 3; ++axis) {
        if (optimum == optimums[axis]) {
            sortctx c = {B->keys[axis]};
            std_sort(&c, B->perm + first, B->perm + last);
            cut = cuts[axis];
            break;
        }
    }
    uint32_t l = init_tree(B, first, cut);
    B->s->nodes[cur_pos].left = l;
    uint32_t r = init_tree(B, cut, last);
    B->s->nodes[cur_pos].right = r;
    return cur_pos;
}

/* bvh.cpp:185-225 */
static rayisec_t bvh_intersect(const orc_scene* s, ray_t ray, float closest, uint32_t v) {
    rayisec_t none; memset(&none, 0, sizeof none); none.id = -1;
    const node_t* nd = &s->nodes[v];
    isec_t bi;
    if (!aabb_intersect(nd->box, ray, &bi)) return none;
    if (closest < bi.t && !bi.interior) return none;
    rayisec_t best; memset(&best, 0, sizeof best); best.isec.t = INF_; best.id = -1;
    if (nd->left == (uint32_t)-1) {
        for (uint32_t i = nd->first; i < nd->first + nd->count; ++i) {
            isec_t is;
            if (prim_intersect(&s->prims[i], ray, &is) && is.t < best.isec.t) { best.isec = is; best.id = (int)i; }
        }
        return best;
    }
    rayisec_t l = bvh_intersect(s, ray, closest, nd->left);
    if (l.id != -1 && l.isec.t < best.isec.t) { closest = l.isec.t; best = l; }
    rayisec_t r = bvh_intersect(s, ray, closest, nd->right);
    if (r.id != -1 && r.isec.t < best.isec.t) best = r;
    return best;
}
/* scene.cpp:46-77 */
/* hw2/hw3 src/scene.cpp:189-211 (closest_dist = -1, `isec.t <= tmax`), hw4 src/scene.cpp:209-227
   (closest_dist = 2e9): every primitive, in file order; the first of equal distances wins */
static rayisec_t ray_intersection_linear(const orc_scene* s, ray_t ray, float tmax) {
    rayisec_t ret; memset(&ret, 0, sizeof ret); ret.id = -1;
    const float plane_tmax = s->dialect <= 3 ? 3.0e38f : 1e5f;
    float closest = s->dialect == 4 ? 2e9f : -1.f;
    for (int i = 0; i < s->nprims; ++i) {
        isec_t is;
        if (!prim_intersect_ex(&s->prims[i], ray, plane_tmax, &is)) continue;
        if (s->dialect == 4) { if (is.t < closest) { closest = is.t; ret.isec = is; ret.id = i; } }
        else if (is.t <= tmax && (closest == -1.f || is.t < closest)) { closest = is.t; ret.isec = is; ret.id = i; }
    }
    return ret;
}
static rayisec_t ray_intersection(const orc_scene* s, ray_t ray) {
    if (s->dialect != 5) return ray_intersection_linear(s, ray, 1e18f);
    rayisec_t ret; memset(&ret, 0, sizeof ret); ret.id = -1;
    float closest = INF_;
    for (int i = 0; i < s->nprims; ++i) {
        if (s->prims[i].type != PT_PLANE) continue;
        isec_t is;
        if (prim_intersect(&s->prims[i], ray, &is) && is.t < closest) { closest = is.t; ret.isec = is; ret.id = i; }
    }
    if (s->nbvh > 0) {
        rayisec_t b = bvh_intersect(s, ray, closest, s->root);
        if (b.id != -1 && b.isec.t < closest) ret = b;
    }
    return ret;
}

/* ------------------------------------------------------------------ scene loading (sceneload.cpp) */
enum { C_EMPTY, C_DIM, C_BG, C_CPOS, C_CRIGHT, C_CUP, C_CFWD, C_FOV, C_NEWPRIM, C_PLANE, C_ELLIPSOID, C_BOX,
       C_POSITION, C_ROTATION, C_COLOR, C_RAYDEPTH, C_METALLIC, C_DIELECTRIC, C_IOR, C_SAMPLES, C_EMISSION,
       C_TRIANGLE, C_AMBIENT, C_NEWLIGHT, C_LINTENS, C_LDIR, C_LPOS, C_LATT, C_UNKNOWN };
/* get_command of each snapshot (hw5 sceneload.cpp:8-33; hw1..hw4 src/scene.cpp:8-36): a word outside the
   snapshot's vocabulary is unknown there */
static int get_command(const char* w, int dialect) {
    static const char* names[] = {"", "DIMENSIONS", "BG_COLOR", "CAMERA_POSITION", "CAMERA_RIGHT", "CAMERA_UP",
        "CAMERA_FORWARD", "CAMERA_FOV_X", "NEW_PRIMITIVE", "PLANE", "ELLIPSOID", "BOX", "POSITION", "ROTATION",
        "COLOR", "RAY_DEPTH", "METALLIC", "DIELECTRIC", "IOR", "SAMPLES", "EMISSION", "TRIANGLE",
        "AMBIENT_LIGHT", "NEW_LIGHT", "LIGHT_INTENSITY", "LIGHT_DIRECTION", "LIGHT_POSITION", "LIGHT_ATTENUATION"};
    static const int first[] = {1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 5, 2, 2, 2, 2, 2, 2};
    static const int last[]  = {5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 5, 2, 2, 2, 2, 2, 2};
    for (int i = 0; i < 28; ++i)
        if (strcmp(w, names[i]) == 0) return (first[i] <= dialect && dialect <= last[i]) ? i : C_UNKNOWN;
    return C_UNKNOWN;
}
typedef struct { const char* p; const char* end; } cursor;
static int next_line(cursor* c, char* buf, size_t cap) {
    if (c->p >= c->end) return 0;
    const char* e = memchr(c->p, '\n', (size_t)(c->end - c->p));
    size_t len = e ? (size_t)(e - c->p) : (size_t)(c->end - c->p);
    size_t k = len < cap - 1 ? len : cap - 1;
    memcpy(buf, c->p, k); buf[k] = 0;
    c->p = e ? e + 1 : c->end;
    return 1;
}
/* operator>> on a stringstream: whitespace-separated tokens; a failed extraction leaves
   later extractions failing too (failbit sticks). */
typedef struct { char* p; int fail; } toks;
static void tok_word(toks* t, char* out, size_t cap) {
    out[0] = 0;
    if (t->fail) return;
    while (*t->p == ' ' || *t->p == '\t' || *t->p == '\r' || *t->p == '\f' || *t->p == '\v') t->p++;
    size_t k = 0;
    while (*t->p && !(*t->p == ' ' || *t->p == '\t' || *t->p == '\r' || *t->p == '\f' || *t->p == '\v')) {
        if (k + 1 < cap) out[k++] = *t->p;
        t->p++;
    }
    out[k] = 0;
    if (k == 0) t->fail = 1;
}
static void tok_float(toks* t, float* out) {
    if (t->fail) return;
    char* e; float v = strtof(t->p, &e);
    if (e == t->p) { t->fail = 1; *out = 0; return; }
    t->p = e; *out = v;
}
static void tok_uint(toks* t, unsigned* out) {
    if (t->fail) return;
    char* e; unsigned long v = strtoul(t->p, &e, 10);
    if (e == t->p) { t->fail = 1; *out = 0; return; }
    t->p = e; *out = (unsigned)v;
}
static void tok_v3(toks* t, v3* v) { tok_float(t, &v->x); tok_float(t, &v->y); tok_float(t, &v->z); }

static void prim_reset(prim_t* p, int type) {
    memset(p, 0, sizeof *p);
    p->type = type; p->rot.w = 1.f; p->material = MAT_DIFFUSE; /* primitives.h:43-47 defaults */
}
/* sceneload.cpp:35-110 ; returns 1 and fills rest[] when an unknown command ended the block */
static int load_primitive(cursor* c, int dialect, prim_t* pr, char* rest, size_t restcap) {
    char line[4096], word[64];
    prim_reset(pr, 0);
    rest[0] = 0;
    while (next_line(c, line, sizeof line)) {
        toks t = {line, 0};
        tok_word(&t, word, sizeof word);
        t.fail = 0;
        int cmd = get_command(word, dialect);
        if (cmd == C_EMPTY) break;
        switch (cmd) {
            case C_ELLIPSOID: { v3 r = {0, 0, 0}; tok_v3(&t, &r); prim_reset(pr, PT_ELLIPSOID); pr->d0 = r; break; }
            case C_PLANE: { v3 n = {0, 0, 0}; tok_v3(&t, &n); prim_reset(pr, PT_PLANE); pr->d0 = n; break; }
            case C_BOX: { v3 s = {0, 0, 0}; tok_v3(&t, &s); prim_reset(pr, PT_BOX); pr->d0 = s; break; }
            case C_TRIANGLE: {
                v3 a = {0, 0, 0}, b = {0, 0, 0}, cc = {0, 0, 0};
                tok_v3(&t, &a); tok_v3(&t, &b); tok_v3(&t, &cc);
                prim_reset(pr, PT_TRIANGLE); pr->d0 = a; pr->d1 = b; pr->d2 = cc; break;
            }
            case C_COLOR: tok_v3(&t, &pr->col); break;
            case C_POSITION: tok_v3(&t, &pr->pos); break;
            case C_ROTATION: tok_float(&t, &pr->rot.x); tok_float(&t, &pr->rot.y); tok_float(&t, &pr->rot.z); tok_float(&t, &pr->rot.w); break;
            case C_METALLIC: pr->material = MAT_METALLIC; break;
            case C_DIELECTRIC: pr->material = MAT_DIELECTRIC; break;
            case C_IOR: tok_float(&t, &pr->ior); break;
            case C_EMISSION: tok_v3(&t, &pr->emission); break;
            default: snprintf(rest, restcap, "%s", word); return 1;
        }
    }
    return 0;
}
static void add_prim(orc_scene* s, const prim_t* p) {
    if (s->nprims == s->cap) { s->cap = s->cap ? s->cap * 2 : 1024; s->prims = (prim_t*)realloc(s->prims, sizeof(prim_t) * (size_t)s->cap); }
    s->prims[s->nprims] = *p; s->prims[s->nprims].orig = s->nprims; s->nprims++;
}
/* hw2 LoadLight, hw2 src/scene.cpp:120-168 */
static int load_light(cursor* c, plight_t* l, char* rest, size_t restcap) {
    char line[4096], word[64];
    memset(l, 0, sizeof *l);
    rest[0] = 0;
    while (next_line(c, line, sizeof line)) {
        toks t = {line, 0};
        tok_word(&t, word, sizeof word);
        t.fail = 0;
        int cmd = get_command(word, 2);
        if (cmd == C_EMPTY) break;
        switch (cmd) {
            case C_LINTENS: tok_v3(&t, &l->intensity); break;
            case C_LPOS: tok_v3(&t, &l->pos); break;
            case C_LDIR: tok_v3(&t, &l->dir); l->directed = 1; break;
            case C_LATT: tok_v3(&t, &l->att); break;
            default: snprintf(rest, restcap, "%s", word); return 1;
        }
    }
    return 0;
}
static void init_scene(orc_scene* s);
static void init_scene_linear(orc_scene* s);

orc_scene* orc_scene_parse(const char* text, long len) { return orc_scene_parse_dialect(text, len, 5); }
orc_scene* orc_scene_parse_dialect(const char* text, long len, int dialect) { /* sceneload.cpp:112-176; hwN src/scene.cpp Scene::Load */
    orc_scene* s = (orc_scene*)calloc(1, sizeof *s);
    s->dialect = dialect;
    cursor c = {text, text + len};
    char line[4096], word[64], rest[64];
    while (next_line(&c, line, sizeof line)) {
        toks t = {line, 0};
        tok_word(&t, word, sizeof word);
        t.fail = 0;
        for (;;) { /* "pasrse_command_again" */
            int cmd = get_command(word, dialect);
            int again = 0;
            switch (cmd) {
                case C_EMPTY: break;
                case C_DIM: tok_uint(&t, &s->width); tok_uint(&t, &s->height); break;
                case C_BG: tok_v3(&t, &s->bg); break;
                case C_CPOS: tok_v3(&t, &s->cam_pos); break;
                case C_CRIGHT: tok_v3(&t, &s->cam_right); break;
                case C_CUP: tok_v3(&t, &s->cam_up); break;
                case C_CFWD: tok_v3(&t, &s->cam_forward); break;
                case C_FOV: tok_float(&t, &s->fov_x); break;
                case C_RAYDEPTH: tok_uint(&t, &s->ray_depth); break;
                case C_SAMPLES: tok_uint(&t, &s->samples); break;
                case C_AMBIENT: tok_v3(&t, &s->ambient); break;
                case C_NEWLIGHT:
                case C_NEWPRIM: {
                    prim_t p;
                    int has_rest;
                    if (cmd == C_NEWLIGHT) {
                        plight_t l;
                        has_rest = load_light(&c, &l, rest, sizeof rest);
                        if (s->nplights == s->plcap) { s->plcap = s->plcap ? s->plcap * 2 : 8; s->plights = (plight_t*)realloc(s->plights, sizeof(plight_t) * (size_t)s->plcap); }
                        s->plights[s->nplights++] = l;
                    } else {
                        has_rest = load_primitive(&c, dialect, &p, rest, sizeof rest);
                        add_prim(s, &p);
                    }
                    if (has_rest && rest[0]) {
                        /* the reference re-dispatches the leftover command name but keeps the
                           (exhausted) stream of the NEW_PRIMITIVE line: its arguments are lost */
                        snprintf(word, sizeof word, "%s", rest);
                        static char emptyline[1] = {0};
                        t.p = emptyline; t.fail = 1;
                        again = 1;
                    }
                    break;
                }
                default: fprintf(stderr, "unexpected command(%s)\n", word); break;
            }
            if (!again) break;
        }
    }
    if (dialect == 5) init_scene(s);
    else init_scene_linear(s);
    return s;
}
orc_scene* orc_scene_load(const char* path) { return orc_scene_load_dialect(path, 5); }
orc_scene* orc_scene_load_dialect(const char* path, int dialect) {
    FILE* f = fopen(path, "rb");
    if (!f) return NULL;
    fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    char* buf = (char*)malloc((size_t)n + 1);
    if (fread(buf, 1, (size_t)n, f) != (size_t)n) { fclose(f); free(buf); return NULL; }
    fclose(f); buf[n] = 0;
    orc_scene* s = orc_scene_parse_dialect(buf, n, dialect);
    free(buf);
    return s;
}
void orc_scene_free(orc_scene* s) { if (!s) return; free(s->prims); free(s->nodes); free(s->lights); free(s->plights); free(s); }

/* hw1..hw4 keep the primitives in file order and have no BVH (Scene::RayIntersection loops over all of
   them); hw4's InitDistribution is the same list of emissive boxes / ellipsoids (hw4 src/scene.cpp Scene::Load tail) */
static void init_scene_linear(orc_scene* s) {
    int n = s->nprims;
    s->nbvh = 0;
    s->lights = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    s->nlights = 0;
    if (s->dialect != 4) return;
    for (int i = 0; i < n; ++i) {
        const prim_t* p = &s->prims[i];
        if (!(p->emission.x > 0 || p->emission.y > 0 || p->emission.z > 0)) continue;
        if (p->type == PT_BOX || p->type == PT_ELLIPSOID) s->lights[s->nlights++] = i;
    }
}

/* scene.cpp:7-40 : InitBVH (partition non-planes first, build) + InitDistribution */
static void init_scene(orc_scene* s) {
    int n = s->nprims;
    int32_t* perm = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
    uint8_t* pred = (uint8_t*)malloc((size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; ++i) { perm[i] = i; pred[i] = s->prims[i].type != PT_PLANE; }
    long nb = orc_partition_flags(perm, pred, n);
    s->nbvh = (int)nb;
    prim_t* src = (prim_t*)malloc(sizeof(prim_t) * (size_t)(n > 0 ? n : 1));
    memcpy(src, s->prims, sizeof(prim_t) * (size_t)n);
    if (nb > 0) {
        build_t B; B.s = s; B.perm = perm; B.src = src;
        B.pbox = (aabb_t*)malloc(sizeof(aabb_t) * (size_t)n);
        for (int a = 0; a < 3; ++a) B.keys[a] = (float*)malloc(sizeof(float) * (size_t)n);
        for (int i = 0; i < n; ++i) {
            if (src[i].type != PT_PLANE) B.pbox[i] = aabb_of_prim(&src[i]); else B.pbox[i] = aabb_empty();
            B.keys[0][i] = src[i].pos.x; B.keys[1][i] = src[i].pos.y; B.keys[2][i] = src[i].pos.z;
        }
        B.cut_qual = (float*)calloc((size_t)nb + 1, sizeof(float));
        s->root = init_tree(&B, 0, (uint32_t)nb); /* bvh.cpp:99-103 */
        free(B.pbox); free(B.cut_qual);
        for (int a = 0; a < 3; ++a) free(B.keys[a]);
    }
    for (int i = 0; i < n; ++i) s->prims[i] = src[perm[i]];
    free(src); free(perm); free(pred);
    /* InitDistribution: emissive boxes / ellipsoids, in final primitive order (scene.cpp:27-40) */
    s->lights = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    s->nlights = 0;
    for (int i = 0; i < n; ++i) {
        const prim_t* p = &s->prims[i];
        if (!(p->emission.x > 0 || p->emission.y > 0 || p->emission.z > 0)) continue;
        if (p->type == PT_BOX || p->type == PT_ELLIPSOID) s->lights[s->nlights++] = i;
    }
}

/* ------------------------------------------------------------------ Philox4x32-10 */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
/* RNG stream contract shared with the CUDA path (DESIGN.md "RNG streams"):
   key=(seed, 0x52544300) ctr=(pixel, sample, slot, block); slot 0 = camera jitter,
   slot b>=1 = shading at the b-th hit.  u01 = (x>>8)*2^-24.
   Box-Muller: r = sqrt(-2 ln(((x0>>8)+1)*2^-24)), th = 2pi*u01(x1): z0 = r cos th, z1 = r sin th */
typedef struct { uint32_t seed, pixel, sample, slot; } rng_t;
static void rng_block(const rng_t* g, uint32_t block, uint32_t out[4]) {
    uint32_t ctr[4] = {g->pixel, g->sample, g->slot, block};
    uint32_t key[2] = {g->seed, 0x52544300u};
    orc_philox4x32_10(ctr, key, out);
}
static inline float u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }
static void box_muller(uint32_t x0, uint32_t x1, float* z0, float* z1) {
    float a = (float)((x0 >> 8) + 1u) * (1.0f / 16777216.0f);
    float r = sqrtf(-2.0f * logf(a));
    float th = 6.283185307179586f * u01(x1);
    *z0 = r * cosf(th); *z1 = r * sinf(th);
}
static v3 normal_vec(const uint32_t b[4]) { /* distributions.cpp:102-110 SampleNormal01Vec */
    float z0, z1, z2, z3;
    box_muller(b[0], b[1], &z0, &z1);
    box_muller(b[2], b[3], &z2, &z3);
    (void)z3;
    return vnormalize(V(z0, z1, z2));
}

/* ------------------------------------------------------------------ distributions.cpp */
static float kPI_(void) { static float k = 0; if (k == 0) k = (float)acos(-1); return k; } /* distributions.h:14 */

/* distributions.cpp:144-159 */
static v3 sample_cosine(const rng_t* g, v3 n) {
    uint32_t b[4]; rng_block(g, 1, b);
    v3 dir = vadd(normal_vec(b), n);
    if (vdot(dir, n) <= 1e-8f) return n;
    if (vlength(dir) <= 1e-4) return n;
    return vnormalize(dir);
}
/* distributions.cpp:161-164 */
static float pdf_cosine(v3 n, v3 d) { return fmaxstd(0.f, 1.f / kPI_() * vdot(d, n)); }

#define REJECT_CAP 64 /* the reference loops forever; both oracle and CUDA path stop after 64 tries */
/* distributions.cpp:227-269 */
static v3 sample_box(const orc_scene* s, const prim_t* box, const rng_t* g, v3 x) {
    (void)s;
    float sx = box->d0.x, sy = box->d0.y, sz = box->d0.z;
    float wx = sx * sx, wy = sy * sy, wz = sz * sz;
    v3 smp = V(0, 0, 0);
    for (int j = 0; j < REJECT_CAP; ++j) {
        uint32_t A[4], Bk[4];
        rng_block(g, 2 + 2 * (uint32_t)j, A); rng_block(g, 3 + 2 * (uint32_t)j, Bk);
        float u = u01(A[0]);
        float side = (u01(A[1]) <= 0.5 ? 1 : -1);
        u *= wx + wy + wz;
        float c1 = u01(A[2]), c2 = u01(A[3]), c3 = u01(Bk[0]);
        c1 = 2 * c1 - 1; c2 = 2 * c2 - 1; c3 = 2 * c3 - 1;
        v3 pnt = V(c1 * sx, c2 * sy, c3 * sz);
        if (u < wx) pnt.x = side * sx;
        else if (u < wx + wy) pnt.y = side * sy;
        else pnt.z = side * sz;
        v3 on_box = vadd(qrot(box->rot, pnt), box->pos);
        smp = vnormalize(vsub(on_box, x));
        ray_t r = {x, smp}; isec_t is;
        if (prim_intersect(box, r, &is)) break;
    }
    return smp;
}
/* distributions.cpp:318-338 */
static v3 sample_ellipsoid(const prim_t* el, const rng_t* g, v3 x) {
    v3 smp = V(0, 0, 0);
    for (int j = 0; j < REJECT_CAP; ++j) {
        uint32_t A[4]; rng_block(g, 2 + 2 * (uint32_t)j, A);
        v3 k = normal_vec(A);
        v3 pnt = vmul(el->d0, k);
        v3 on = vadd(qrot(el->rot, pnt), el->pos);
        smp = vnormalize(vsub(on, x));
        ray_t r = {x, smp}; isec_t is;
        if (prim_intersect(el, r, &is)) break;
    }
    return smp;
}
/* distributions.cpp:385-399 */
static v3 mix_sample(const orc_scene* s, const rng_t* g, v3 x, v3 n) {
    uint32_t b0[4]; rng_block(g, 0, b0);
    float flip = u01(b0[0]);
    if (s->nlights == 0 || flip <= 0.5f) return sample_cosine(g, n);
    float fid = u01(b0[1]);
    size_t id = (size_t)floorf(fid * (float)s->nlights);
    const prim_t* l = &s->prims[s->lights[id]];
    if (l->type == PT_BOX) return sample_box(s, l, g, x);
    return sample_ellipsoid(l, g, x);
}
/* distributions.cpp:170-198 ; returns number of points found (0,1,2) */
static int points_for_pdf(const prim_t* pr, v3 x, v3 d, isec_t* i1, isec_t* i2) {
    ray_t r = {x, d};
    if (!prim_intersect(pr, r, i1)) return 0;
    float t = i1->t;
    if (t <= 1e-8) return 0;
    float e = 1e-4f;
    v3 inner = vadd(x, vscale(t + e, d));
    ray_t r2 = {inner, d};
    if (!prim_intersect(pr, r2, i2)) return 1;
    i2->t += t + e;
    return 2;
}
/* distributions.cpp:271-287 */
static float pdf_point_box(const prim_t* box, float dist2, v3 n, v3 d) {
    float sx = box->d0.x, sy = box->d0.y, sz = box->d0.z;
    float wx = sx * sx, wy = sy * sy, wz = sz * sz;
    float p_y = (float)(1. / (double)(2 * 4 * (wx + wy + wz)));
    return p_y * dist2 / fabsf(vdot(d, n));
}
/* distributions.cpp:340-347 */
static float pdf_point_ellipsoid(const prim_t* el, float dist2, v3 y, v3 n_, v3 d) {
    v3 r = el->d0;
    v3 n = vdiv(qrot(qconj(el->rot), vsub(y, el->pos)), r);
    float len = vlength(V(n.x * r.y * r.z, r.x * n.y * r.z, r.x * r.y * n.z));
    float p_y = (float)(1. / (double)(4 * kPI_() * len));
    return p_y * dist2 / fabsf(vdot(d, n_));
}
static float dist2(v3 a, v3 b) { v3 q = vsub(b, a); return vdot(q, q); } /* glm::distance2 */
/* distributions.cpp:289-312 and :349-372 */
static float pdf_light(const prim_t* pr, v3 x, v3 d) {
    isec_t i1, i2;
    int k = points_for_pdf(pr, x, d, &i1, &i2);
    if (k == 0) return 1e-9f;
    v3 p1 = vadd(x, vscale(i1.t, d));
    float sum = pr->type == PT_BOX ? pdf_point_box(pr, dist2(p1, x), i1.n, d)
                                   : pdf_point_ellipsoid(pr, dist2(p1, x), p1, i1.n, d);
    if (k == 2) {
        v3 p2 = vadd(x, vscale(i2.t, d));
        sum += pr->type == PT_BOX ? pdf_point_box(pr, dist2(p2, x), i2.n, d)
                                  : pdf_point_ellipsoid(pr, dist2(p2, x), p2, i2.n, d);
    }
    return sum;
}
/* distributions.cpp:401-416 */
static float mix_pdf(const orc_scene* s, v3 x, v3 n, v3 d) {
    float sum = pdf_cosine(n, d);
    if (s->nlights > 0) {
        float prim_sum = 0.f;
        for (int i = 0; i < s->nlights; ++i) prim_sum += pdf_light(&s->prims[s->lights[i]], x, d);
        prim_sum *= 1.f / (float)s->nlights;
        sum = 0.5f * sum + 0.5f * prim_sum;
    }
    return sum;
}

/* ------------------------------------------------------------------ camera / integrator */
/* scene.cpp:180-187 */
static ray_t cam_ray(const orc_scene* s, float x, float y) {
    float tan_fov_x = (float)tan((double)(s->fov_x / 2));
    float tan_fov_y = tan_fov_x * (float)s->height / (float)s->width;
    float nx = (2 * x / (float)s->width - 1) * tan_fov_x;
    float ny = -1.f * (2 * y / (float)s->height - 1) * tan_fov_y;
    ray_t r;
    r.o = s->cam_pos;
    r.d = vadd(vadd(vscale(nx, s->cam_right), vscale(ny, s->cam_up)), vscale(1.f, s->cam_forward));
    return r;
}
static v3 reflect_dir(v3 normal, v3 dir) { /* scene.cpp:79-81 */
    v3 two_n = vscale(2.0f, normal);
    float k = vdot(normal, dir);
    return vsub(dir, V(two_n.x * k, two_n.y * k, two_n.z * k));
}
/* scene.cpp:83-178 unrolled into a loop: L = sum_k beta_k * E_k, beta_k = prod of bounce weights */
static v3 ray_trace(const orc_scene* s, uint32_t seed, uint32_t pixel, uint32_t sample, ray_t ray, uint64_t* nrays) {
    v3 L = V(0, 0, 0), beta = V(1, 1, 1);
    const float SCENE_EPS = s->dialect <= 3 ? 1e-3f : 1e-4f; /* hw2/hw3 include/scene.h:60, hw4/hw5 include/scene.h:57/64 */
    for (unsigned bounce = 1; bounce <= s->ray_depth; ++bounce) {
        rayisec_t h = ray_intersection(s, ray);
        (*nrays)++;
        if (h.id == -1) { L = vadd(L, vmul(beta, s->bg)); break; }
        const prim_t* pr = &s->prims[h.id];
        float t = h.isec.t; v3 normal = h.isec.n; int interior = h.isec.interior;
        v3 p = vadd(ray.o, vscale(t, ray.d));
        L = vadd(L, vmul(beta, pr->emission));
        rng_t g = {seed, pixel, sample, bounce};
        if (pr->material == MAT_DIFFUSE && s->dialect == 3) {
            /* hw3 src/scene.cpp:238-249: a direction uniform on the sphere, mirrored into the hemisphere of
               the normal; L = E + C * 2 * dot(w, n) * L_in(w) */
            uint32_t b1[4]; rng_block(&g, 1, b1);
            v3 dir = normal_vec(b1);
            float cs = vdot(dir, normal);
            if (cs < 0) { dir = vscale(-1.f, dir); cs = -cs; }
            beta = vmul(beta, V(pr->col.x * (2 * cs), pr->col.y * (2 * cs), pr->col.z * (2 * cs)));
            ray.o = vadd(p, vscale(SCENE_EPS, dir)); ray.d = dir;
        } else if (pr->material == MAT_DIFFUSE) {
            v3 p_outer = vadd(p, vscale(SCENE_EPS, normal));
            v3 dir = mix_sample(s, &g, p_outer, normal);
            float cs = vdot(dir, normal);
            if (cs <= 0) break;
            float pw = mix_pdf(s, p_outer, normal, dir);
            /* C / kPI: glm vec/scalar divides per component */
            v3 w = V(pr->col.x / kPI_(), pr->col.y / kPI_(), pr->col.z / kPI_());
            float k2 = 1 / pw;
            beta = vmul(beta, V(w.x * cs * k2, w.y * cs * k2, w.z * cs * k2));
            ray.o = vadd(p, vscale(SCENE_EPS, dir)); ray.d = dir;
        } else if (pr->material == MAT_METALLIC) {
            v3 rd = reflect_dir(normal, vnormalize(ray.d));
            beta = vmul(beta, pr->col);
            ray.o = vadd(p, vscale(SCENE_EPS, rd)); ray.d = rd;
        } else { /* DIELECTRIC scene.cpp:130-170 */
            float eta1 = 1.f, eta2 = pr->ior;
            if (interior) { float tmp = eta1; eta1 = eta2; eta2 = tmp; }
            v3 dir = vscale(-1.f, vnormalize(ray.d));
            float dn = vdot(normal, dir);
            float sin2 = (float)((double)(eta1 / eta2) * sqrt((double)fmaxstd(0.f, 1 - dn * dn))); /* ::sqrt(double) */
            v3 rd = reflect_dir(normal, vnormalize(ray.d));
            int reflect = 0;
            if (fabsf(sin2) > 1.) reflect = 1;
            else {
                float r0 = (float)pow((double)((eta1 - eta2) / (eta1 + eta2)), 2.);
                float r = (float)((double)r0 + (double)(1 - r0) * pow((double)(1 - dn), 5.));
                uint32_t b0[4]; rng_block(&g, 0, b0);
                if (u01(b0[0]) < r) reflect = 1;
            }
            if (reflect) { ray.o = vadd(p, vscale(SCENE_EPS, rd)); ray.d = rd; }
            else {
                float cos2 = sqrtf(1 - sin2 * sin2);
                float e = eta1 / eta2;
                v3 a = vscale(e, vscale(-1.f, dir));
                v3 b = vscale(e * dn - cos2, normal);
                v3 fr = vadd(a, b);
                ray.o = vadd(p, vscale(SCENE_EPS, fr)); ray.d = fr;
                if (!interior) beta = vmul(beta, pr->col);
            }
        }
    }
    return L;
}

/* ---- hw2: Scene::RayTrace, hw2 src/scene.cpp:262-341 (recursive, both dielectric branches followed) */
static void calc_light(const plight_t* l, v3 p, v3* colour, v3* dir, float* dist) { /* hw2 src/lights.cpp:7-23 */
    if (l->directed) { *colour = l->intensity; *dir = vnormalize(l->dir); *dist = 1e18f; return; }
    v3 to = vsub(l->pos, p);
    float d = vlength(to);
    float k = (float)(1. / (double)(l->att.x + l->att.y * d + l->att.z * d * d));
    *colour = vscale(k, l->intensity); *dir = vnormalize(to); *dist = d;
}
static v3 whitted(const orc_scene* s, ray_t ray, unsigned depth, uint64_t* nrays) {
    const float eps = 1e-3f; /* hw2 include/scene.h:60 */
    if (depth == 0) return V(0, 0, 0);
    rayisec_t h = ray_intersection_linear(s, ray, 1e18f);
    (*nrays)++;
    if (h.id == -1) return s->bg;
    const prim_t* pr = &s->prims[h.id];
    v3 normal = h.isec.n; int interior = h.isec.interior;
    v3 p = vadd(ray.o, vscale(h.isec.t, ray.d));
    v3 rd = reflect_dir(normal, vnormalize(ray.d));
    if (pr->material == MAT_DIFFUSE) {
        v3 sum = s->ambient;
        for (int l = 0; l < s->nplights; ++l) {
            v3 colour, dir; float dist;
            calc_light(&s->plights[l], p, &colour, &dir, &dist);
            float k = vdot(dir, normal);
            if (k >= 0) {
                ray_t sh = {vadd(p, vscale(eps, dir)), dir};
                (*nrays)++;
                if (ray_intersection_linear(s, sh, dist).id == -1) sum = vadd(sum, vscale(k, colour));
            }
        }
        return vmul(sum, pr->col);
    }
    ray_t rr = {vadd(p, vscale(eps, rd)), rd};
    if (pr->material == MAT_METALLIC) return vmul(pr->col, whitted(s, rr, depth - 1, nrays));
    v3 reflected = whitted(s, rr, depth - 1, nrays);
    float eta1 = 1.f, eta2 = pr->ior;
    if (interior) { float tmp = eta1; eta1 = eta2; eta2 = tmp; }
    v3 dir = vscale(-1.f, vnormalize(ray.d));
    float dn = vdot(normal, dir);
    float sin2 = (float)((double)(eta1 / eta2) * sqrt((double)(1 - dn * dn)));
    if (fabsf(sin2) > 1.) return reflected;
    float cos2 = (float)sqrt((double)(1 - sin2 * sin2));
    float e = eta1 / eta2;
    v3 fr = vadd(vscale(e, vscale(-1.f, dir)), vscale(e * dn - cos2, normal));
    ray_t fray = {vadd(p, vscale(eps, fr)), fr};
    v3 refr = whitted(s, fray, depth - 1, nrays);
    if (!interior) refr = vmul(refr, pr->col);
    float r0 = (float)pow((double)((eta1 - eta2) / (eta1 + eta2)), 2.);
    float r = (float)((double)r0 + (double)(1 - r0) * pow((double)(1 - dn), 5.));
    return vadd(vscale(r, reflected), vscale(1 - r, refr));
}
/* pixel-centre camera ray, hw2 src/scene.cpp:229-237: nx, ny evaluated in double with (x + 0.5) */
static ray_t cam_ray_centre(const orc_scene* s, unsigned x, unsigned y) {
    float tan_fov_x = (float)tan((double)(s->fov_x / 2));
    float tan_fov_y = tan_fov_x * (float)s->height / (float)s->width;
    float nx = (float)((2 * (x + 0.5) / s->width - 1) * (double)tan_fov_x);
    float ny = (float)(-1.f * (2 * (y + 0.5) / s->height - 1) * (double)tan_fov_y);
    ray_t r;
    r.o = s->cam_pos;
    r.d = vadd(vadd(vscale(nx, s->cam_right), vscale(ny, s->cam_up)), vscale(1.f, s->cam_forward));
    return r;
}
/* ---- hw1: Scene::raytrace, hw1 src/scene.cpp:40-56, everything in double (hw1 include/point.h) */
typedef struct { double x, y, z; } d3;
static d3 D3(v3 a) { d3 r = {a.x, a.y, a.z}; return r; }
static d3 dsub(d3 a, d3 b) { d3 r = {a.x - b.x, a.y - b.y, a.z - b.z}; return r; }
static double ddot(d3 a, d3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static d3 dcross_hw1(d3 a, d3 p) { d3 r = {a.z * p.y - a.y * p.z, a.x * p.z - a.z * p.x, a.y * p.x - a.x * p.y}; return r; } /* point.cpp:21-23 */
typedef struct { d3 v; double w; } dq;
static dq dqmul(dq a, dq q) { /* quaternion.cpp:11-13 */
    d3 c = dcross_hw1(a.v, q.v);
    dq r = {{a.w * q.v.x + q.w * a.v.x + c.x, a.w * q.v.y + q.w * a.v.y + c.y, a.w * q.v.z + q.w * a.v.z + c.z}, a.w * q.w - ddot(a.v, q.v)};
    return r;
}
static d3 dqtransform(dq q, d3 x) { /* quaternion.cpp:19-21 */
    dq px = {x, 0.}, cj = {{-q.v.x, -q.v.y, -q.v.z}, q.w};
    return dqmul(dqmul(q, px), cj).v;
}
static int hw1_intersect(const prim_t* pr, d3 o, d3 d, double* t_out) { /* hw1 src/primitives.cpp:4-103 */
    dq q = {{pr->rot.x, pr->rot.y, pr->rot.z}, pr->rot.w};
    o = dqtransform(q, dsub(o, D3(pr->pos)));
    d = dqtransform(q, d);
    d3 g = D3(pr->d0);
    if (pr->type == PT_PLANE) {
        double t = -ddot(o, g) / ddot(d, g);
        if (t < 0) return 0;
        *t_out = t; return 1;
    }
    if (pr->type == PT_BOX) {
        double ax = (-g.x - o.x) / d.x, bx = (g.x - o.x) / d.x, ay = (-g.y - o.y) / d.y, by = (g.y - o.y) / d.y;
        double az = (-g.z - o.z) / d.z, bz = (g.z - o.z) / d.z;
        double t1 = fmax(fmax(fmin(ax, bx), fmin(ay, by)), fmin(az, bz));
        double t2 = fmin(fmin(fmax(ax, bx), fmax(ay, by)), fmax(az, bz));
        if (t1 > t2 || t2 < 0) return 0;
        *t_out = t1 < 0 ? t2 : t1; return 1;
    }
    d3 dr = {d.x / g.x, d.y / g.y, d.z / g.z}, orr = {o.x / g.x, o.y / g.y, o.z / g.z};
    double a = ddot(dr, dr), b = 2 * ddot(orr, dr), c = ddot(orr, orr) - 1;
    double disc = b * b - 4 * a * c;
    if (disc <= 0) return 0;
    double x1 = (-b - sqrt(disc)) / (2 * a), x2 = (-b + sqrt(disc)) / (2 * a);
    if (x1 > x2) { double tmp = x1; x1 = x2; x2 = tmp; }
    if (x2 < 0) return 0;
    *t_out = x1 < 0 ? x2 : x1; return 1;
}
static v3 raycast_hw1(const orc_scene* s, unsigned x, unsigned y) {
    double tan_fov_x = tan((double)s->fov_x / 2), tan_fov_y = tan_fov_x * s->height / s->width;
    float nx = (float)((2 * (x + 0.5) / s->width - 1) * tan_fov_x);
    float ny = (float)(-1.f * (2 * (y + 0.5) / s->height - 1) * tan_fov_y);
    d3 r = D3(s->cam_right), u = D3(s->cam_up), f = D3(s->cam_forward);
    d3 d = {nx * r.x + ny * u.x + 1.f * f.x, nx * r.y + ny * u.y + 1.f * f.y, nx * r.z + ny * u.z + 1.f * f.z};
    v3 ans = s->bg;
    double closest = -1;
    for (int i = 0; i < s->nprims; ++i) {
        double t;
        if (hw1_intersect(&s->prims[i], D3(s->cam_pos), d, &t) && (closest == -1 || t < closest)) { closest = t; ans = s->prims[i].col; }
    }
    return ans;
}

void orc_render_sum(const orc_scene* s, uint32_t seed, uint32_t sample_begin, uint32_t sample_count,
                    long pix_begin, long pix_end, float* out_sum, uint64_t counters[2], int nthreads) {
    uint64_t paths = 0, rays = 0;
    if (s->dialect <= 2) { /* deterministic snapshots: the frame itself, once, whatever the sample range */
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : paths, rays)
        for (long i = pix_begin; i < pix_end; ++i) {
            unsigned x = (unsigned)(i % s->width), y = (unsigned)(i / s->width);
            uint64_t nr = 0;
            v3 c = s->dialect == 1 ? raycast_hw1(s, x, y) : whitted(s, cam_ray_centre(s, x, y), s->ray_depth, &nr);
            out_sum[3 * (i - pix_begin) + 0] = c.x; out_sum[3 * (i - pix_begin) + 1] = c.y; out_sum[3 * (i - pix_begin) + 2] = c.z;
            paths++; rays += nr;
        }
        if (counters) { counters[0] += paths; counters[1] += rays; }
        return;
    }
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#else
    (void)nthreads;
#endif
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : paths, rays)
    for (long i = pix_begin; i < pix_end; ++i) {
        unsigned x = (unsigned)(i % s->width), y = (unsigned)(i / s->width);
        v3 sum = V(0, 0, 0);
        for (uint32_t k = 0; k < sample_count; ++k) {
            uint32_t smp = sample_begin + k;
            rng_t g = {seed, (uint32_t)i, smp, 0};
            uint32_t b[4]; rng_block(&g, 0, b);
            float fx = (float)x + u01(b[0]), fy = (float)y + u01(b[1]); /* scene.cpp:197-198 */
            /* hw3's Camera::GetToRay(float, float) still adds the half pixel of its integer ancestor
               (hw3 src/scene.cpp:186-187): the jittered samples cover [x + 0.5, x + 1.5) */
            if (s->dialect == 3) { fx = (float)((double)fx + 0.5); fy = (float)((double)fy + 0.5); }
            uint64_t nr = 0;
            v3 c = ray_trace(s, seed, (uint32_t)i, smp, cam_ray(s, fx, fy), &nr);
            sum = vadd(sum, c);
            paths++; rays += nr;
        }
        out_sum[3 * (i - pix_begin) + 0] = sum.x;
        out_sum[3 * (i - pix_begin) + 1] = sum.y;
        out_sum[3 * (i - pix_begin) + 2] = sum.z;
    }
    if (counters) { counters[0] += paths; counters[1] += rays; }
}

/* ------------------------------------------------------------------ color.cpp */
static float saturate1(float c) { return fmaxstd(fminstd(1.f, c), 0.f); } /* color.cpp:19-24 */
void orc_tonemap_u8(long npix, const float* rgb, uint8_t* out) {
    const float a = 2.51f, b = 0.03f, c = 2.43f, d = 0.59f, e = 0.14f; /* color.cpp:26-35 */
    float gamma = (float)(1. / 2.2);                                   /* color.cpp:37-41 */
    for (long i = 0; i < 3 * npix; ++i) {
        float x = rgb[i];
        float y = saturate1((x * (a * x + b)) / (x * (c * x + d) + e));
        float g = powf(y, gamma);
        out[i] = (unsigned char)round((double)(255 * g)); /* color.cpp:43-49 */
    }
}

/* hw1 Color::toUInts without tone mapping, hw1 src/color.cpp:10-16 */
void orc_flat_u8(long npix, const float* rgb, uint8_t* out) {
    for (long i = 0; i < 3 * npix; ++i) out[i] = (unsigned char)round(255 * (double)rgb[i]);
}
int orc_scene_dialect(const orc_scene* s) { return s->dialect; }

/* ------------------------------------------------------------------ batch entry points */
void orc_scene_info(const orc_scene* s, uint32_t out[8]) {
    out[0] = s->width; out[1] = s->height; out[2] = s->ray_depth; out[3] = s->samples;
    out[4] = (uint32_t)s->nprims; out[5] = (uint32_t)s->nbvh; out[6] = (uint32_t)s->nnodes; out[7] = (uint32_t)s->nlights;
}
void orc_scene_camera(const orc_scene* s, float o[16]) {
    const v3* src[4] = {&s->cam_pos, &s->cam_right, &s->cam_up, &s->cam_forward};
    for (int i = 0; i < 4; ++i) { o[3 * i] = src[i]->x; o[3 * i + 1] = src[i]->y; o[3 * i + 2] = src[i]->z; }
    o[12] = s->fov_x; o[13] = s->bg.x; o[14] = s->bg.y; o[15] = s->bg.z;
}
void orc_scene_override(orc_scene* s, int width, int height, int samples, int ray_depth) {
    if (width >= 0) s->width = (unsigned)width;
    if (height >= 0) s->height = (unsigned)height;
    if (samples >= 0) s->samples = (unsigned)samples;
    if (ray_depth >= 0) s->ray_depth = (unsigned)ray_depth;
}
void orc_scene_prim_order(const orc_scene* s, int32_t* out) { for (int i = 0; i < s->nprims; ++i) out[i] = s->prims[i].orig; }
void orc_scene_prims(const orc_scene* s, int32_t* tm, float* d) {
    for (int i = 0; i < s->nprims; ++i) {
        const prim_t* p = &s->prims[i];
        tm[2 * i] = p->type; tm[2 * i + 1] = p->material;
        float* o = d + 26 * (long)i;
        o[0] = p->col.x; o[1] = p->col.y; o[2] = p->col.z;
        o[3] = p->emission.x; o[4] = p->emission.y; o[5] = p->emission.z;
        o[6] = p->pos.x; o[7] = p->pos.y; o[8] = p->pos.z;
        o[9] = p->rot.x; o[10] = p->rot.y; o[11] = p->rot.z; o[12] = p->rot.w;
        o[13] = p->ior;
        o[14] = p->d0.x; o[15] = p->d0.y; o[16] = p->d0.z;
        o[17] = p->d1.x; o[18] = p->d1.y; o[19] = p->d1.z;
        o[20] = p->d2.x; o[21] = p->d2.y; o[22] = p->d2.z;
        o[23] = 0; o[24] = 0; o[25] = 0;
    }
}
void orc_scene_nodes(const orc_scene* s, float* aabb, uint32_t* links) {
    for (int i = 0; i < s->nnodes; ++i) {
        const node_t* n = &s->nodes[i];
        aabb[6 * i] = n->box.mn.x; aabb[6 * i + 1] = n->box.mn.y; aabb[6 * i + 2] = n->box.mn.z;
        aabb[6 * i + 3] = n->box.mx.x; aabb[6 * i + 4] = n->box.mx.y; aabb[6 * i + 5] = n->box.mx.z;
        links[4 * i] = n->left; links[4 * i + 1] = n->right; links[4 * i + 2] = n->first; links[4 * i + 3] = n->count;
    }
}
uint32_t orc_scene_root(const orc_scene* s) { return s->root; }

void orc_intersect(const orc_scene* s, long n, const float* o, const float* d,
                   int32_t* id, float* t, float* normal, int32_t* interior) {
#pragma omp parallel for schedule(dynamic, 256)
    for (long i = 0; i < n; ++i) {
        ray_t r = {V(o[3 * i], o[3 * i + 1], o[3 * i + 2]), V(d[3 * i], d[3 * i + 1], d[3 * i + 2])};
        rayisec_t h = ray_intersection(s, r);
        id[i] = h.id;
        t[i] = h.id == -1 ? 0.f : h.isec.t;
        normal[3 * i] = h.id == -1 ? 0.f : h.isec.n.x;
        normal[3 * i + 1] = h.id == -1 ? 0.f : h.isec.n.y;
        normal[3 * i + 2] = h.id == -1 ? 0.f : h.isec.n.z;
        interior[i] = h.id == -1 ? 0 : h.isec.interior;
    }
}
void orc_primitive_intersect(const orc_scene* s, int prim, long n, const float* o, const float* d,
                             int32_t* hit, float* t, float* normal, int32_t* interior) {
    for (long i = 0; i < n; ++i) {
        ray_t r = {V(o[3 * i], o[3 * i + 1], o[3 * i + 2]), V(d[3 * i], d[3 * i + 1], d[3 * i + 2])};
        isec_t is; memset(&is, 0, sizeof is);
        hit[i] = prim_intersect(&s->prims[prim], r, &is);
        t[i] = hit[i] ? is.t : 0.f;
        normal[3 * i] = hit[i] ? is.n.x : 0.f; normal[3 * i + 1] = hit[i] ? is.n.y : 0.f; normal[3 * i + 2] = hit[i] ? is.n.z : 0.f;
        interior[i] = hit[i] ? is.interior : 0;
    }
}
void orc_camera_rays(const orc_scene* s, long n, const float* xy, float* o, float* d) {
    for (long i = 0; i < n; ++i) {
        ray_t r = cam_ray(s, xy[2 * i], xy[2 * i + 1]);
        o[3 * i] = r.o.x; o[3 * i + 1] = r.o.y; o[3 * i + 2] = r.o.z;
        d[3 * i] = r.d.x; d[3 * i + 1] = r.d.y; d[3 * i + 2] = r.d.z;
    }
}
void orc_mix_pdf(const orc_scene* s, long n, const float* x, const float* nrm, const float* d, float* pdf) {
    for (long i = 0; i < n; ++i)
        pdf[i] = mix_pdf(s, V(x[3 * i], x[3 * i + 1], x[3 * i + 2]), V(nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]),
                         V(d[3 * i], d[3 * i + 1], d[3 * i + 2]));
}
void orc_mix_sample(const orc_scene* s, long n, const float* x, const float* nrm,
                    uint32_t seed, uint32_t sample, uint32_t bounce, float* dir) {
    for (long i = 0; i < n; ++i) {
        rng_t g = {seed, (uint32_t)i, sample, bounce};
        v3 r = mix_sample(s, &g, V(x[3 * i], x[3 * i + 1], x[3 * i + 2]), V(nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]));
        dir[3 * i] = r.x; dir[3 * i + 1] = r.y; dir[3 * i + 2] = r.z;
    }
}
