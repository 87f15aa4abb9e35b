// helper_math.h — stand-in for the NVIDIA CUDA Samples header of that name (volumeRender_kernel.cu:17; not
// vendored by the reference).  The component-wise vector arithmetic the reference's device code uses, written
// from the operations' definitions; normalize(v) = v * rsqrtf(dot(v, v)) as in the samples.  Test infrastructure.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#define VRDD_HD inline __host__ __device__

VRDD_HD float3 make_float3(float s) { return make_float3(s, s, s); }
VRDD_HD float3 make_float3(float4 a) { return make_float3(a.x, a.y, a.z); }
VRDD_HD float4 make_float4(float s) { return make_float4(s, s, s, s); }
VRDD_HD float4 make_float4(float3 a, float w) { return make_float4(a.x, a.y, a.z, w); }

VRDD_HD float3 operator-(float3 a) { return make_float3(-a.x, -a.y, -a.z); }
VRDD_HD float3 operator+(float3 a, float3 b) { return make_float3(a.x + b.x, a.y + b.y, a.z + b.z); }
VRDD_HD float3 operator-(float3 a, float3 b) { return make_float3(a.x - b.x, a.y - b.y, a.z - b.z); }
VRDD_HD float3 operator*(float3 a, float3 b) { return make_float3(a.x * b.x, a.y * b.y, a.z * b.z); }
VRDD_HD float3 operator/(float3 a, float3 b) { return make_float3(a.x / b.x, a.y / b.y, a.z / b.z); }
VRDD_HD float3 operator+(float3 a, float b) { return make_float3(a.x + b, a.y + b, a.z + b); }
VRDD_HD float3 operator-(float3 a, float b) { return make_float3(a.x - b, a.y - b, a.z - b); }
VRDD_HD float3 operator*(float3 a, float b) { return make_float3(a.x * b, a.y * b, a.z * b); }
VRDD_HD float3 operator*(float b, float3 a) { return make_float3(b * a.x, b * a.y, b * a.z); }
VRDD_HD float3 operator/(float3 a, float b) { return make_float3(a.x / b, a.y / b, a.z / b); }
VRDD_HD float3 operator/(float b, float3 a) { return make_float3(b / a.x, b / a.y, b / a.z); }
VRDD_HD void operator+=(float3& a, float3 b) { a.x += b.x; a.y += b.y; a.z += b.z; }
VRDD_HD void operator-=(float3& a, float3 b) { a.x -= b.x; a.y -= b.y; a.z -= b.z; }
VRDD_HD void operator*=(float3& a, float b) { a.x *= b; a.y *= b; a.z *= b; }

VRDD_HD float4 operator+(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
VRDD_HD float4 operator-(float4 a, float4 b) { return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
VRDD_HD float4 operator*(float4 a, float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
VRDD_HD float4 operator*(float4 a, float b) { return make_float4(a.x * b, a.y * b, a.z * b, a.w * b); }
VRDD_HD float4 operator*(float b, float4 a) { return make_float4(b * a.x, b * a.y, b * a.z, b * a.w); }
VRDD_HD float4 operator/(float4 a, float b) { return make_float4(a.x / b, a.y / b, a.z / b, a.w / b); }
VRDD_HD void operator+=(float4& a, float4 b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
VRDD_HD void operator-=(float4& a, float4 b) { a.x -= b.x; a.y -= b.y; a.z -= b.z; a.w -= b.w; }
VRDD_HD void operator*=(float4& a, float b) { a.x *= b; a.y *= b; a.z *= b; a.w *= b; }

VRDD_HD float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
VRDD_HD float dot(float4 a, float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
VRDD_HD float3 fminf(float3 a, float3 b) { return make_float3(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z)); }
VRDD_HD float3 fmaxf(float3 a, float3 b) { return make_float3(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)); }
VRDD_HD float3 normalize(float3 v) {
#ifdef __CUDA_ARCH__
    const float inv_len = rsqrtf(dot(v, v));
#else
    const float inv_len = 1.0f / sqrtf(dot(v, v));
#endif
    return v * inv_len;
}
VRDD_HD float length(float3 v) { return sqrtf(dot(v, v)); }
VRDD_HD float clamp(float f, float a, float b) { return fmaxf(a, fminf(f, b)); }
VRDD_HD float lerp(float a, float b, float t) { return a + t * (b - a); }
#undef VRDD_HD
