// device_scene.h -- what a kernel sees of a scene: pointers into HBM plus a few scalars, passed
// BY VALUE as a kernel parameter (lives in the constant bank, no extra load to reach a pointer).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rtc {

struct DevScene {
    // per primitive (final order), 16-byte aligned float4 arrays
    const float4* geo0;    // triangle (a, n.x) | other (d0, 0)
    const float4* geo1;    // triangle (b, n.y)
    const float4* geo2;    // triangle (c, n.z)
    const float4* xf_pos;  // (pos, bits(type | flags))
    const float4* xf_rot;  // quaternion xyzw
    const float4* mat0;    // (colour, bits(material))
    const float4* mat1;    // (emission, ior)
    // index BVH: 4 float4 (64 B) per 4-wide node, fp16 child boxes
    const float4* inodes;
    // reference BVH: 2 float4 per node + meta
    const float4* rnodes;
    const uint4* rmeta;
    const uint32_t* lca;
    const int32_t* lights;
    const float4* ubox;    // per primitive slot: exact (min) (max) of the reference leaf that starts there
    const float4* planes;  // per plane: (n.xyz, bits(prim id)) (pos.xyz, bits(1 = no rotation))

    uint32_t nprims, nbvh, nnodes, root, iroot, lca_levels, nlights, ref_depth;
    uint32_t width, height, ray_depth, nplanes;
    float3 cam_pos, cam_right, cam_up, cam_forward;
    float tan_fov_x, tan_fov_y;
    float3 bg;
};

// wavefront path state, structure of float4 arrays
struct PathSoA {
    float4* o;      // origin.xyz, -
    float4* d;      // direction.xyz, -
    float4* beta;   // throughput.rgb, bits(pixel index)
    float4* rad;    // radiance so far .rgb, bits(sample index)
};
struct HitSoA {
    float* cd;      // distance of the closest plane (1e18 = none): closest_dist handed to the BVH
    uint32_t* id;   // 0xFFFFFFFF = miss, else id of the closest primitive
};

constexpr uint32_t HIT_MISS = 0xFFFFFFFFu;
constexpr uint32_t HIT_INTERIOR = 1u << 30;

}  // namespace rtc
