// svr_math.cuh -- float3 arithmetic and the arithmetic class of the render path.
//
// The reference is built with -use_fast_math (CMakeLists.txt:9) on top of GLM vectors; this library
// is compiled with the same flag so logf/expf/powf/sinf/cosf, division, rsqrt and FTZ fall in the
// same intrinsic class.  Vector helpers follow GLM's definitions where the definition is visible
// in results: normalize(v) = v * (1/sqrt(dot(v,v))), reflect(I,N) = I - N*dot(N,I)*2,
// min/max = `a < b ? a : b` (not NaN-swallowing), cross as GLM writes it.
#pragma once

#include <cuda_runtime.h>
#include <cfloat>
#include <cstdint>

#include "../../include/svr_types.h"

#define SVR_DEV __device__ __forceinline__
#define SVR_HD __host__ __device__ __forceinline__

#define SVR_PI_F 3.14159265358979323846f
#define SVR_INV_PI_F 0.31830988618379067154f
#define SVR_PI_D 3.14159265358979323846
#define SVR_INV_PI_D 0.31830988618379067154

namespace svr {

SVR_HD float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
SVR_HD float3 f3(float s) { return make_float3(s, s, s); }
SVR_HD float3 f3(const svr_vec3& v) { return make_float3(v.x, v.y, v.z); }

SVR_HD float3 operator+(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
SVR_HD float3 operator-(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
SVR_HD float3 operator-(float3 a) { return f3(-a.x, -a.y, -a.z); }
SVR_HD float3 operator*(float3 a, float3 b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
SVR_HD float3 operator*(float3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
SVR_HD float3 operator*(float s, float3 a) { return f3(s * a.x, s * a.y, s * a.z); }
SVR_HD float3 operator/(float3 a, float s) { return f3(a.x / s, a.y / s, a.z / s); }
SVR_HD float3 operator/(float s, float3 a) { return f3(s / a.x, s / a.y, s / a.z); }
SVR_HD float3& operator+=(float3& a, float3 b) { a = a + b; return a; }
SVR_HD float3& operator*=(float3& a, float3 b) { a = a * b; return a; }

SVR_HD float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
SVR_HD float3 cross(float3 a, float3 b)
{
    return f3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
}
SVR_DEV float3 normalize(float3 a) { return a * rsqrtf(dot(a, a)); }
SVR_DEV float3 reflect(float3 i, float3 n) { return i - n * (dot(n, i) * 2.f); }
SVR_HD float max3(float3 a) { return fmaxf(a.x, fmaxf(a.y, a.z)); }
SVR_HD float min3(float3 a) { return fminf(a.x, fminf(a.y, a.z)); }

// Orthonormal basis from one vector, core/cuda_onb.h:26-40.
struct Onb {
    float3 u, v, w;
    SVR_DEV explicit Onb(float3 w_)
    {
        w = w_;
        if (fabsf(w.x) > fabsf(w.y)) {
            float inv = rsqrtf(w.x * w.x + w.z * w.z);
            v = f3(-w.z * inv, 0.f, w.x * inv);
        } else {
            float inv = rsqrtf(w.y * w.y + w.z * w.z);
            v = f3(0.f, w.z * inv, -w.y * inv);
        }
        u = cross(v, w);
    }
    SVR_DEV float3 local(float a, float b, float c) const { return a * u + b * v + c * w; }
};

struct Ray {
    float3 orig, dir;
};

}  // namespace svr
