// TEST INFRASTRUCTURE -- not product code.
//
// Minimal stand-in for the subset of GLM (0.9.6-0.9.8 era, un-vendored third-party dependency
// of the reference, CMakeLists.txt:37) that the reference's render hot path uses.  It exists only
// so that oracle/build_ref.sh can compile the reference's pathtracer.cu / raycasting.cu
// *unmodified, from where they lie* into oracle/_ref/.  The product (sunvolumerender_b200/csrc)
// never includes this file; it has its own float3 math.
//
// Usage census that defines the subset: vec2/vec3/vec4/ivec3/u8vec4,
// dot/cross/normalize/reflect/min/max/length, the `glm::uninitialize` ctor tag, `.a`/`.w`
// aliasing on vec4, vec2::operator[], component-wise vec*vec, scalar/vec.  Layouts are the packed
// GLM ones (vec3 = 12 B align 4, u8vec4 = 4 B align 1) because they are part of the struct ABI the
// boundary passes by value.
#pragma once

#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define SVR_GLM_FN __host__ __device__ inline
#else
#define SVR_GLM_FN inline
#endif

namespace glm {

enum ctor { uninitialize };

template <typename T>
struct tvec2 {
    T x, y;
    SVR_GLM_FN tvec2() {}
    SVR_GLM_FN explicit tvec2(ctor) {}
    SVR_GLM_FN explicit tvec2(T s) : x(s), y(s) {}
    SVR_GLM_FN tvec2(T a, T b) : x(a), y(b) {}
    SVR_GLM_FN T& operator[](int i) { return (&x)[i]; }
    SVR_GLM_FN const T& operator[](int i) const { return (&x)[i]; }
};

template <typename T>
struct tvec4;

template <typename T>
struct tvec3 {
    T x, y, z;
    SVR_GLM_FN tvec3() {}
    SVR_GLM_FN explicit tvec3(ctor) {}
    SVR_GLM_FN explicit tvec3(T s) : x(s), y(s), z(s) {}
    template <typename A, typename B, typename C>
    SVR_GLM_FN tvec3(A a, B b, C c) : x(static_cast<T>(a)), y(static_cast<T>(b)), z(static_cast<T>(c)) {}
    SVR_GLM_FN explicit tvec3(const tvec4<T>& v);
    SVR_GLM_FN T& operator[](int i) { return (&x)[i]; }
    SVR_GLM_FN const T& operator[](int i) const { return (&x)[i]; }
    SVR_GLM_FN tvec3& operator+=(const tvec3& o) { x += o.x; y += o.y; z += o.z; return *this; }
    SVR_GLM_FN tvec3& operator-=(const tvec3& o) { x -= o.x; y -= o.y; z -= o.z; return *this; }
    SVR_GLM_FN tvec3& operator*=(const tvec3& o) { x *= o.x; y *= o.y; z *= o.z; return *this; }
    SVR_GLM_FN tvec3& operator/=(const tvec3& o) { x /= o.x; y /= o.y; z /= o.z; return *this; }
    SVR_GLM_FN tvec3& operator*=(T s) { x *= s; y *= s; z *= s; return *this; }
    SVR_GLM_FN tvec3& operator/=(T s) { x /= s; y /= s; z /= s; return *this; }
};

template <typename T>
struct tvec4 {
    union { T x; T r; };
    union { T y; T g; };
    union { T z; T b; };
    union { T w; T a; };
    SVR_GLM_FN tvec4() {}
    SVR_GLM_FN explicit tvec4(ctor) {}
    SVR_GLM_FN explicit tvec4(T s) : x(s), y(s), z(s), w(s) {}
    template <typename A, typename B, typename C, typename D>
    SVR_GLM_FN tvec4(A a_, B b_, C c_, D d_)
        : x(static_cast<T>(a_)), y(static_cast<T>(b_)), z(static_cast<T>(c_)), w(static_cast<T>(d_)) {}
    SVR_GLM_FN tvec4& operator+=(const tvec4& o) { x += o.x; y += o.y; z += o.z; w += o.w; return *this; }
    SVR_GLM_FN tvec4& operator*=(T s) { x *= s; y *= s; z *= s; w *= s; return *this; }
};

template <typename T>
SVR_GLM_FN tvec3<T>::tvec3(const tvec4<T>& v) : x(v.x), y(v.y), z(v.z) {}

typedef tvec2<float> vec2;
typedef tvec3<float> vec3;
typedef tvec4<float> vec4;
typedef tvec3<int> ivec3;
// -DSVR_REF_FLOAT_TWIN: the reference kernels end in `img[offset] = glm::u8vec4(L.x * 255, ...)`
// (raycasting.cu:66, pathtracer.cu:289), which truncates to 8 bits and makes a "within 1e-4" check
// meaningless.  With this switch the SAME unmodified kernel source stores those four products as
// floats (16 bytes per pixel), giving a float twin of the reference without touching its code.
#ifdef SVR_REF_FLOAT_TWIN
typedef tvec4<float> u8vec4;
#else
typedef tvec4<uint8_t> u8vec4;
#endif

// ---- vec2 ----
SVR_GLM_FN vec2 operator+(const vec2& a, const vec2& b) { return vec2(a.x + b.x, a.y + b.y); }
SVR_GLM_FN vec2 operator-(const vec2& a, const vec2& b) { return vec2(a.x - b.x, a.y - b.y); }
SVR_GLM_FN vec2 operator*(const vec2& a, float s) { return vec2(a.x * s, a.y * s); }
SVR_GLM_FN vec2 operator*(float s, const vec2& a) { return vec2(a.x * s, a.y * s); }

// ---- vec3 ----
SVR_GLM_FN vec3 operator-(const vec3& a) { return vec3(-a.x, -a.y, -a.z); }
SVR_GLM_FN vec3 operator+(const vec3& a, const vec3& b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
SVR_GLM_FN vec3 operator-(const vec3& a, const vec3& b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
SVR_GLM_FN vec3 operator*(const vec3& a, const vec3& b) { return vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
SVR_GLM_FN vec3 operator/(const vec3& a, const vec3& b) { return vec3(a.x / b.x, a.y / b.y, a.z / b.z); }
SVR_GLM_FN vec3 operator*(const vec3& a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
SVR_GLM_FN vec3 operator*(float s, const vec3& a) { return vec3(s * a.x, s * a.y, s * a.z); }
SVR_GLM_FN vec3 operator/(const vec3& a, float s) { return vec3(a.x / s, a.y / s, a.z / s); }
SVR_GLM_FN vec3 operator/(float s, const vec3& a) { return vec3(s / a.x, s / a.y, s / a.z); }
SVR_GLM_FN vec3 operator+(const vec3& a, float s) { return vec3(a.x + s, a.y + s, a.z + s); }
SVR_GLM_FN vec3 operator-(const vec3& a, float s) { return vec3(a.x - s, a.y - s, a.z - s); }

// ---- vec4 ----
SVR_GLM_FN vec4 operator+(const vec4& a, const vec4& b) { return vec4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
SVR_GLM_FN vec4 operator*(const vec4& a, float s) { return vec4(a.x * s, a.y * s, a.z * s, a.w * s); }
SVR_GLM_FN vec4 operator*(float s, const vec4& a) { return vec4(s * a.x, s * a.y, s * a.z, s * a.w); }

// ---- geometric ----
SVR_GLM_FN float dot(const vec2& a, const vec2& b) { return a.x * b.x + a.y * b.y; }
SVR_GLM_FN float dot(const vec3& a, const vec3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
SVR_GLM_FN float length(const vec3& a) { return sqrtf(dot(a, a)); }
SVR_GLM_FN vec3 cross(const vec3& a, const vec3& b)
{
    return vec3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
}
// GLM: normalize(v) = v * inversesqrt(dot(v, v))
SVR_GLM_FN vec3 normalize(const vec3& a) { return a * (1.f / sqrtf(dot(a, a))); }
// GLM: reflect(I, N) = I - N * dot(N, I) * 2
SVR_GLM_FN vec3 reflect(const vec3& i, const vec3& n) { return i - n * (dot(n, i) * 2.f); }
// GLM: min(x, y) = x < y ? x : y, max(x, y) = x > y ? x : y (component-wise; not NaN-swallowing)
SVR_GLM_FN vec3 min(const vec3& a, const vec3& b) { return vec3(a.x < b.x ? a.x : b.x, a.y < b.y ? a.y : b.y, a.z < b.z ? a.z : b.z); }
SVR_GLM_FN vec3 max(const vec3& a, const vec3& b) { return vec3(a.x > b.x ? a.x : b.x, a.y > b.y ? a.y : b.y, a.z > b.z ? a.z : b.z); }

}  // namespace glm
