// dev_math.cuh — device-side fp64/fp32 helpers of the tracing kernels.
// Compiled with -fmad=false: every expression keeps the reference's operation order and rounding
// (ray_tracer.cu:72-139, triangle_mesh.cu:39-137, normal_shader.cu:48-124), so that results are
// bit-comparable with a host evaluation of the same source expressions.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

struct d3 { double x, y, z; };
struct f3 { float x, y, z; };

__device__ __forceinline__ d3 mk3(double x, double y, double z) { d3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ d3 operator+(d3 a, d3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ d3 operator-(d3 a, d3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ d3 operator*(double a, d3 b) { return mk3(a * b.x, a * b.y, a * b.z); }
__device__ __forceinline__ d3 cross3(d3 a, d3 b) { return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
__device__ __forceinline__ double dot3(d3 a, d3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ double magsq3(d3 a) { return (a.x * a.x + a.y * a.y + a.z * a.z); }
__device__ __forceinline__ double len3(d3 a) { return sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }
// Three IEEE divisions by one divisor (a vector normalisation) from ONE correctly rounded reciprocal y = RN(1/n):
// q0 = RN(a y), r = a - q0 n (exact, FMA), q = RN(q0 + r y) is the correctly rounded a / n (Markstein's final correction
// step — the step CUDA's own division ends with); 3 instructions per quotient instead of a division each.  Quotients
// outside the comfortable range (zero, subnormal, huge, n = 0) take the plain division.  tools/markstein_check.c compares
// the sequence with a / n on the host (FMA hardware) for 3e9 vectors / adversarial significand patterns: no mismatch.
__device__ __noinline__ double div_plain(double a, double n) { return a / n; }
__device__ __forceinline__ double div_by(double a, double n, double y)
{
    const double q0 = __dmul_rn(a, y);
    const double q = __fma_rn(__fma_rn(-q0, n, a), y, q0);
    const double m = fabs(q0);
    if (m > 1e-290 && m < 1e290) return q;
    if (a == 0.0 && n > 0.0 && n < 1e290) return q0;        // +-0 / n = +-0 (axis-parallel vectors: common, and not worth a call)
    return div_plain(a, n);
}
// (used by the direction pass only: in the shading code the same sequence costs k_primary_follow more in registers —
// 486 instead of 306 bytes of spills at 72 registers, 1.60 -> 1.67 ms — than the divisions it saves)
__device__ __forceinline__ d3 normalised3_shared_rcp(d3 a)
{
    const double n = len3(a), y = __drcp_rn(n);
    return mk3(div_by(a.x, n, y), div_by(a.y, n, y), div_by(a.z, n, y));
}
__device__ __forceinline__ d3 normalised3(d3 a) { double n = len3(a); return mk3(a.x / n, a.y / n, a.z / n); }
// fp64 normalise, then narrow (ray_tracer.cu:125-129, normal_shader.cu:96-100)
__device__ __forceinline__ f3 normalise_float3(d3 a)
{
    double n = len3(a);
    f3 r; r.x = (float)(a.x / n); r.y = (float)(a.y / n); r.z = (float)(a.z / n);
    return r;
}

// ---- the OptiX 6.x math the reference calls (optixu_math_namespace.h, restated) ----
__device__ __forceinline__ float dotf3(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ f3 normalizef3(f3 v)
{
    float invLen = 1.0f / sqrtf(dotf3(v, v));
    f3 r; r.x = v.x * invLen; r.y = v.y * invLen; r.z = v.z * invLen;
    return r;
}
// reflect(i, n) = i - 2.0f * n * dot(n, i)            (normal_shader.cu:296)
__device__ __forceinline__ f3 optix_reflect(f3 i, f3 n)
{
    float d = dotf3(n, i);
    f3 r; r.x = i.x - (2.0f * n.x) * d; r.y = i.y - (2.0f * n.y) * d; r.z = i.z - (2.0f * n.z) * d;
    return r;
}
// refract(r, i, n, ior), false on total internal reflection  (normal_shader.cu:212)
__device__ __forceinline__ bool optix_refract(f3 &r, f3 i, f3 n, float ior)
{
    f3 nn = n;
    float negNdotV = dotf3(i, nn);
    float eta;
    if (negNdotV > 0.0f) {
        eta = ior;
        nn.x = -n.x; nn.y = -n.y; nn.z = -n.z;
        negNdotV = -negNdotV;
    } else {
        eta = 1.f / ior;
    }
    const float k = 1.f - eta * eta * (1.f - negNdotV * negNdotV);
    if (k < 0.0f) {
        r.x = r.y = r.z = 0.f;
        return false;
    }
    const float s = eta * negNdotV + sqrtf(k);
    f3 v; v.x = eta * i.x - s * nn.x; v.y = eta * i.y - s * nn.y; v.z = eta * i.z - s * nn.z;
    r = normalizef3(v);
    return true;
}

// ray_tracer.cu:53-69
__device__ __forceinline__ void normalise_angle(double &angle)
{
    while (angle < -M_PI) angle += 2 * M_PI;
    while (angle > M_PI) angle -= 2 * M_PI;
}
__device__ __forceinline__ bool angle_in_range(double testAngle, double a, double b)
{
    a -= testAngle;
    b -= testAngle;
    normalise_angle(a);
    normalise_angle(b);
    if (a * b >= 0) return false;
    return fabs(a - b) < M_PI;
}
