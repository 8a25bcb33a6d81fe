// rts_optix_shim.h — a host-side emulation of the small part of the legacy OptiX (<= 6.5) device
// API that the reference's three program files use, so that those files can be compiled
// UNMODIFIED with g++ (from /root/reference, never copied) and executed on the CPU.
// TEST INFRASTRUCTURE: used only by ref_harness.cpp → oracle/_ref/libref_rts.so, which in turn is
// used only by tests/ and tests/golden/make_golden.py to pin the oracle restatement.
//
// What is emulated, and from where:
//   rtDeclareVariable / rtBuffer / RT_PROGRAM / rtPrintf   : trivial host stand-ins
//   rtTrace            : exhaustive search over every triangle of every target (OptiX's BVH is
//                        closed source; exhaustive search is the definitional closest hit)
//   rtPotentialIntersection(t) : tmin < (float)t < current closest (strict, fp32)
//   rtReportIntersection       : commit attributes of the potential hit
//   optix::reflect / refract / normalize / dot / make_Ray / RT_DEFAULT_MAX / Aabb :
//                        restated from memory of optixu_math_namespace.h (OptiX 6.x). These are
//                        the only "parity unpinned" pieces: no OptiX SDK is available offline.
#ifndef RTS_OPTIX_SHIM_H
#define RTS_OPTIX_SHIM_H

#include <vector_types.h>
#include <vector_functions.h>
#include <math.h>
#include <stdio.h>
#include <stddef.h>
#include <string.h>

#undef __device__
#define __device__
#undef __host__
#define __host__
#define RT_PROGRAM
#define RT_DEFAULT_MAX 1.e27f
#define rtPrintf(...) ((void)0)
#define rtDeclareVariable(type, name, semantic, annotation) static type name

typedef int rtObject;

template <typename T, int N> struct rtBuffer;
template <typename T> struct rtBuffer<T, 1> {
    T *data = nullptr;
    size_t n = 0;
    void set(T *p, size_t count) { data = p; n = count; }
    T &operator[](size_t i) { return data[i]; }
    size_t size() const { return n; }
};
template <typename T> struct rtBuffer<T, 2> {
    T *data = nullptr;
    size_t w = 0, h = 0;
    void set(T *p, size_t width, size_t height) { data = p; w = width; h = height; }
    T &operator[](uint2 i) { return data[(size_t)i.x + (size_t)i.y * w]; }
};

namespace optix {

struct Ray {
    float3 origin;
    float3 direction;
    unsigned int ray_type;
    float tmin;
    float tmax;
};
static inline Ray make_Ray(float3 origin, float3 direction, unsigned int ray_type, float tmin, float tmax)
{
    Ray r;
    r.origin = origin; r.direction = direction; r.ray_type = ray_type; r.tmin = tmin; r.tmax = tmax;
    return r;
}

static inline float3 operator-(const float3 &a) { return make_float3(-a.x, -a.y, -a.z); }
static inline float3 operator-(const float3 &a, const float3 &b) { return make_float3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline float3 operator*(const float3 &a, const float3 &b) { return make_float3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline float3 operator*(const float3 &a, const float s) { return make_float3(a.x * s, a.y * s, a.z * s); }
static inline float3 operator*(const float s, const float3 &a) { return make_float3(a.x * s, a.y * s, a.z * s); }
static inline float dot(const float3 &a, const float3 &b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline float3 normalize(const float3 &v)
{
    float invLen = 1.0f / sqrtf(dot(v, v));
    return v * invLen;
}
static inline float3 reflect(const float3 &i, const float3 &n) { return i - 2.0f * n * dot(n, i); }
static inline bool refract(float3 &r, const float3 &i, const float3 &n, const float ior)
{
    float3 nn = n;
    float negNdotV = dot(i, nn);
    float eta;
    if (negNdotV > 0.0f) {
        eta = ior;
        nn = -n;
        negNdotV = -negNdotV;
    } else {
        eta = 1.f / ior;
    }
    const float k = 1.f - eta * eta * (1.f - negNdotV * negNdotV);
    if (k < 0.0f) {
        r = make_float3(0.f, 0.f, 0.f);
        return false;
    } else {
        r = normalize(eta * i - (eta * negNdotV + sqrtf(k)) * nn);
        return true;
    }
}

struct Aabb {
    float3 m_min, m_max;
    void invalidate()
    {
        m_min = make_float3(1e37f, 1e37f, 1e37f);
        m_max = make_float3(-1e37f, -1e37f, -1e37f);
    }
};

} // namespace optix

// CUDA directed-rounding conversions used by the bounding-box program (triangle_mesh.cu:228-229)
static inline float __double2float_rd(double x)
{
    float f = (float)x;
    if ((double)f > x) f = nextafterf(f, -INFINITY);
    return f;
}
static inline float __double2float_ru(double x)
{
    float f = (float)x;
    if ((double)f < x) f = nextafterf(f, INFINITY);
    return f;
}

// traversal entry points implemented by ref_harness.cpp
namespace shim {
void trace(const optix::Ray &ray, void *payload, size_t payload_size);
bool potential_intersection(float t);
bool report_intersection(unsigned int material);
} // namespace shim

template <class P> static inline void rtTrace(rtObject, const optix::Ray &ray, P &prd) { shim::trace(ray, &prd, sizeof(P)); }
static inline bool rtPotentialIntersection(float t) { return shim::potential_intersection(t); }
static inline bool rtReportIntersection(unsigned int m) { return shim::report_intersection(m); }

#endif
