// vecmath.h -- float3 / quaternion helpers shared by the host scene builder and the CUDA kernels.
//
// The operation ORDER of every helper follows glm 1.0.0 as the reference uses it
// (dot = x*x' + y*y' + z*z' summed left to right, normalize = v * (1/sqrt(dot)), quaternion
// rotation = v + ((uv*w) + uuv)*2), so that host-side precomputation (triangle normals, AABBs,
// SAH costs) is bit-identical to what the reference computes when built without FMA
// contraction (csrc/Makefile compiles host code with -ffp-contract=off).
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define RT_HD __host__ __device__ __forceinline__
#else
#define RT_HD inline
#endif

namespace rtc {

struct vec3 {
    float x, y, z;
};
struct quat {
    float x, y, z, w;
};

RT_HD vec3 mk3(float x, float y, float z) { vec3 r; r.x = x; r.y = y; r.z = z; return r; }
RT_HD vec3 operator+(vec3 a, vec3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
RT_HD vec3 operator-(vec3 a, vec3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
RT_HD vec3 operator*(vec3 a, vec3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
RT_HD vec3 operator/(vec3 a, vec3 b) { return mk3(a.x / b.x, a.y / b.y, a.z / b.z); }
RT_HD vec3 operator*(float k, vec3 a) { return mk3(k * a.x, k * a.y, k * a.z); }
RT_HD vec3 operator*(vec3 a, float k) { return mk3(a.x * k, a.y * k, a.z * k); }
RT_HD vec3 operator-(vec3 a) { return mk3(-a.x, -a.y, -a.z); }
RT_HD float dot(vec3 a, vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
RT_HD vec3 cross(vec3 a, vec3 b) {
    return mk3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
}
RT_HD float length(vec3 a) { return sqrtf(dot(a, a)); }
RT_HD vec3 normalize(vec3 a) {
    float k = 1.0f / sqrtf(dot(a, a));
    return mk3(a.x * k, a.y * k, a.z * k);
}
RT_HD quat conjugate(quat q) { quat r; r.x = -q.x; r.y = -q.y; r.z = -q.z; r.w = q.w; return r; }
RT_HD vec3 rotate(quat q, vec3 v) {
    vec3 qv = mk3(q.x, q.y, q.z);
    vec3 uv = cross(qv, v);
    vec3 uuv = cross(qv, uv);
    vec3 s = (uv * q.w) + uuv;
    return v + s * 2.0f;
}
RT_HD float idx(vec3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

}  // namespace rtc
