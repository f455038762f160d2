/*
 * rt_oracle.h -- CPU oracle for the hw5 path tracer hot path (TEST INFRASTRUCTURE ONLY).
 *
 * This is a plain-C restatement of the reference's per-pixel Monte Carlo path
 * (/root/reference/hw5/src/{scene,sceneload,bvh,primitives,distributions,color}.cpp).
 * It exists to CHECK the CUDA path; nothing in the product (raytracing-course_b200/)
 * may include, link or call it.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py use it.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks every deterministic
 * function below against fixtures produced by the UNMODIFIED reference compiled from
 * /root/reference (oracle/refprobe.cpp, oracle/Makefile -> oracle/_ref/), see
 * tests/golden/README.md.  The only intentional deviation is the random number
 * stream: north_star replaces std::minstd_rand by counter-based Philox4x32-10, so
 * rendered images are compared statistically, never bit-wise.
 */
#ifndef RT_ORACLE_H
#define RT_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_scene orc_scene;

/* Scene::Load + Scene::InitScene (sceneload.cpp:110-175, scene.cpp:7-40). NULL on IO error. */
orc_scene* orc_scene_load(const char* path);
orc_scene* orc_scene_parse(const char* text, long len);
void orc_scene_free(orc_scene* s);
/* The four earlier snapshots of the renderer (/root/reference/hw1..hw4/src/scene.cpp): dialect 1..4 selects that
   snapshot's scene vocabulary, its Scene::RayIntersection (a loop over all primitives in file order) and its
   Scene::RayTrace: 1 ray casting in double, 2 Whitted with point / directional lights, 3 path tracing with
   uniform-hemisphere sampling, 4 = hw5's estimator without triangles and BVH.  orc_render_sum on a dialect 1 / 2
   scene returns the deterministic frame (linear colour, once).  Pinned by tests/golden/hw*_*.npz (images of the
   compiled reference programs); known deviation: hw2 / hw3 primitives are rotated with glm's float quaternion-vector
   product as in hw4/hw5, the snapshots themselves use the sandwich product q v q* (identical for unit quaternions,
   scaled by |q|^2 otherwise); hw1 is restated in double with its own sandwich product. */
orc_scene* orc_scene_load_dialect(const char* path, int dialect);
orc_scene* orc_scene_parse_dialect(const char* text, long len, int dialect);
int orc_scene_dialect(const orc_scene* s);
void orc_flat_u8(long npix, const float* rgb_linear, uint8_t* out);

/* header values */
void orc_scene_info(const orc_scene* s, uint32_t out[8]);
/* out: width,height,ray_depth,samples,nprims,nbvh(non-plane prims),nnodes,nlights */
void orc_scene_camera(const orc_scene* s, float out[16]);
/* out: pos(3) right(3) up(3) forward(3) fov_x bg(3) */
void orc_scene_override(orc_scene* s, int width, int height, int samples, int ray_depth); /* <0 keeps */

/* final primitive order after std::partition + BVH sorts: orig index of each slot */
void orc_scene_prim_order(const orc_scene* s, int32_t* out_orig_index);
/* per-primitive record in FINAL order: type,material, then 26 floats
   (col3 emission3 pos3 rot4(xyzw) ior d0(3) d1(3) d2(3)) */
void orc_scene_prims(const orc_scene* s, int32_t* type_material /*2*n*/, float* data /*26*n*/);
/* BVH nodes in reference vector order: per node aabb_min3 aabb_max3 and left,right,first,count */
void orc_scene_nodes(const orc_scene* s, float* aabb /*6*nnodes*/, uint32_t* links /*4*nnodes*/);
uint32_t orc_scene_root(const orc_scene* s);

/* Scene::RayIntersection (scene.cpp:46-77) for n rays. id=-1 on miss. */
void orc_intersect(const orc_scene* s, long n, const float* o, const float* d,
                   int32_t* id, float* t, float* normal, int32_t* interior);
/* Primitive::Intersect (primitives.cpp:14-52) of one primitive (FINAL order id). hit[i]=0/1 */
void orc_primitive_intersect(const orc_scene* s, int prim, long n, const float* o, const float* d,
                             int32_t* hit, float* t, float* normal, int32_t* interior);
/* Camera::GetToRay (scene.cpp:180-187) */
void orc_camera_rays(const orc_scene* s, long n, const float* xy, float* o, float* d);
/* Distribution::Pdf of the scene mix distribution (distributions.cpp:401-416) */
void orc_mix_pdf(const orc_scene* s, long n, const float* x, const float* nrm, const float* d, float* pdf);
/* Distribution::Sample of the mix (distributions.cpp:385-399) driven by Philox stream
   (seed, pixel=idx, sample, bounce). */
void orc_mix_sample(const orc_scene* s, long n, const float* x, const float* nrm,
                    uint32_t seed, uint32_t sample, uint32_t bounce, float* dir);

/* AcesTonemap + GammaCorrected + Color::toUInts (color.cpp:26-49) */
void orc_tonemap_u8(long npix, const float* rgb_linear, uint8_t* out);

/* Scene::Render restated with Philox streams: mean radiance over samples
   [sample_begin, sample_begin+sample_count) for pixels [pix_begin,pix_end), linear float rgb
   SUM (not mean) written to out_sum[3*(pix-pix_begin)].  counters[0]+=paths, [1]+=rays. */
void orc_render_sum(const orc_scene* s, uint32_t seed, uint32_t sample_begin, uint32_t sample_count,
                    long pix_begin, long pix_end, float* out_sum, uint64_t counters[2], int nthreads);

/* Philox4x32-10 block, exposed so tests can pin it against the published KAT. */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

/* libstdc++-order restatements, exposed for pinning against the real std:: in tests */
void orc_sort_perm_by_key(const float* key, int32_t* perm, long first, long last);
long orc_partition_flags(int32_t* perm, const uint8_t* pred_true /*indexed by perm value*/, long n);

#ifdef __cplusplus
}
#endif
#endif
