// gk_device.cuh — device-side vector helpers and the sampling/RNG library of the path tracer.
// Restates assets/shaders/common/Const_Func.slang:4-49, 227-340 and PreProcessor.slang for CUDA.
// Compiled with -fmad=false: every a*b+c below is two rounded operations unless written as fmaf/__fmaf_rn
// (include/gk_detmath.h explains the numeric contract with the CPU checker).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gk_detmath.h"
#include "../../include/gkcu.h"

#define GK_DEV __device__ __forceinline__

namespace gkd {

GK_DEV float3 f3(float a, float b, float c) { return make_float3(a, b, c); }
GK_DEV float4 f4(float a, float b, float c, float d) { return make_float4(a, b, c, d); }
GK_DEV float3 operator+(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
GK_DEV float3 operator-(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
GK_DEV float3 operator-(float3 a) { return f3(-a.x, -a.y, -a.z); }
GK_DEV float3 operator*(float3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
GK_DEV float3 operator*(float3 a, float3 b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
GK_DEV float4 operator+(float4 a, float4 b) { return f4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
GK_DEV float4 operator-(float4 a, float4 b) { return f4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
GK_DEV float4 operator*(float4 a, float s) { return f4(a.x * s, a.y * s, a.z * s, a.w * s); }
GK_DEV float4 operator*(float4 a, float4 b) { return f4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
GK_DEV float4 operator/(float4 a, float s) { return f4(a.x / s, a.y / s, a.z / s, a.w / s); }
GK_DEV float2 operator+(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
GK_DEV float2 operator-(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
GK_DEV float2 operator*(float2 a, float s) { return make_float2(a.x * s, a.y * s); }
GK_DEV float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
GK_DEV float3 cross(float3 a, float3 b) { return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
GK_DEV float length(float3 a) { return sqrtf(dot(a, a)); }
GK_DEV float length(float2 a) { return sqrtf(a.x * a.x + a.y * a.y); }
GK_DEV float3 normalize(float3 a) { float inv = 1.0f / sqrtf(dot(a, a)); return a * inv; }
GK_DEV float3 xyz(float4 a) { return f3(a.x, a.y, a.z); }
GK_DEV float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }
GK_DEV float saturatef(float v) { return clampf(v, 0.0f, 1.0f); }
GK_DEV float3 lerp3(float3 a, float3 b, float t) { return a + (b - a) * t; }
GK_DEV float4 lerp4(float4 a, float4 b, float t) { return a + (b - a) * t; }

// mul(M, v), glm column-major m[col][row]
GK_DEV float4 mul(const GkMat4& M, float4 v) {
    float4 r;
    r.x = M.m[0][0] * v.x + M.m[1][0] * v.y + M.m[2][0] * v.z + M.m[3][0] * v.w;
    r.y = M.m[0][1] * v.x + M.m[1][1] * v.y + M.m[2][1] * v.z + M.m[3][1] * v.w;
    r.z = M.m[0][2] * v.x + M.m[1][2] * v.y + M.m[2][2] * v.z + M.m[3][2] * v.w;
    r.w = M.m[0][3] * v.x + M.m[1][3] * v.y + M.m[2][3] * v.z + M.m[3][3] * v.w;
    return r;
}

// fp16x4 storage images (R16G16B16A16_SFLOAT): round-to-nearest-even stores, exact loads; 8-byte vector access.
GK_DEV float4 load_h4(const uint2* img, size_t i) {
    uint2 v = img[i];
    __half2 a = *reinterpret_cast<__half2*>(&v.x), b = *reinterpret_cast<__half2*>(&v.y);
    float2 fa = __half22float2(a), fb = __half22float2(b);
    return f4(fa.x, fa.y, fb.x, fb.y);
}
GK_DEV uint2 pack_h4(float4 c) {
    __half2 a = __floats2half2_rn(c.x, c.y), b = __floats2half2_rn(c.z, c.w);
    uint2 v;
    v.x = *reinterpret_cast<uint32_t*>(&a);
    v.y = *reinterpret_cast<uint32_t*>(&b);
    return v;
}
GK_DEV void store_h4(uint2* img, size_t i, float4 c) { img[i] = pack_h4(c); }
GK_DEV float h2f(uint16_t h) { return __half2float(__ushort_as_half(h)); }

// ---- PreProcessor.slang
#define GK_M_PI 3.14159265358979323846f
#define GK_M_TWO_PI 6.283185307179586476925f
#define GK_M_PI_4 0.785398163397448309616f
#define GK_M_1_PI 0.318309886183790671538f
#define GK_M_1_OVER_TWO_PI 0.159154943091895335768f
#define GK_EPS 1e-3f
#define GK_NEARZERO 1e-35f
#define GK_PT_MAX_TRACE_DISTANCE 1000.f
#define GK_PRIMARY_T_MAX 2000.f
#define GK_COS_QUARTER_DEGREE 0.99999048f

GK_DEV bool isZERO(float x) { return x > -GK_NEARZERO && x < GK_NEARZERO; }
GK_DEV float3 to_local(float3 v, float3 t, float3 b, float3 n) { return f3(dot(t, v), dot(b, v), dot(n, v)); }
GK_DEV float3 to_world(float3 v, float3 T, float3 B, float3 N) {
    return f3(v.x * T.x + v.y * B.x + v.z * N.x, v.x * T.y + v.y * B.y + v.z * N.y, v.x * T.z + v.y * B.z + v.z * N.z);
}

// ---- Const_Func.slang
GK_DEV float pow5(float x) { return x * x * x * x * x; }
GK_DEV float Schlick(float cosine, float refractionIndex) {  // :8-14
    float r0 = (1 - refractionIndex) / (1 + refractionIndex);
    r0 *= r0;
    return r0 + (1 - r0) * pow5(1 - cosine);
}
GK_DEV void ONB(float3 n, float3& b1, float3& b2) {  // :17-22
    float signZ = n.z < 0.f ? -1.f : 1.f;
    float a = -1.0f / (signZ + n.z);
    b2 = f3(n.x * n.y * a, signZ + n.y * n.y * a, -n.y);
    b1 = f3(1.0f + signZ * n.x * n.x * a, signZ * b2.x, -signZ * n.x);
}
GK_DEV float3 AlignWithNormal(float3 ray, float3 normal) {  // :44-49
    float3 T, B;
    ONB(normal, T, B);
    return to_world(ray, T, B, normal);
}
GK_DEV void pcg4d(uint4& v) {  // :227-233
    v.x = v.x * 1664525u + 1013904223u; v.y = v.y * 1664525u + 1013904223u;
    v.z = v.z * 1664525u + 1013904223u; v.w = v.w * 1664525u + 1013904223u;
    v.x += v.y * v.w; v.y += v.z * v.x; v.z += v.x * v.y; v.w += v.y * v.z;
    v.x ^= v.x >> 16u; v.y ^= v.y >> 16u; v.z ^= v.z >> 16u; v.w ^= v.w >> 16u;
    v.x += v.y * v.w; v.y += v.z * v.x; v.z += v.x * v.y; v.w += v.y * v.z;
}
GK_DEV float uint_to_float(uint32_t x) { return __uint_as_float(0x3f800000u | (x >> 9)) - 1.0f; }  // :235
GK_DEV float RandomFloat(uint4& v) { pcg4d(v); return uint_to_float(v.x); }
GK_DEV float2 RandomFloat2(uint4& v) { pcg4d(v); return make_float2(uint_to_float(v.x), uint_to_float(v.y)); }
GK_DEV float2 concentric_sample_disk(float2 offset) {  // :260-275
    offset = offset + (offset - make_float2(1, 1));
    if (isZERO(offset.x) && isZERO(offset.y)) return make_float2(0, 0);
    if (fabsf(offset.x) > fabsf(offset.y)) {
        float theta = GK_M_PI_4 * offset.y / offset.x;
        return make_float2(gk_det_cosf(theta), gk_det_sinf(theta)) * offset.x;
    }
    float cos_theta = gk_det_sinf(GK_M_PI_4 * offset.x / offset.y);
    return make_float2(cos_theta, sqrtf(1.f - cos_theta * cos_theta)) * offset.y;
}
GK_DEV float2 RandomInUnitDisk(uint4& seed) { return concentric_sample_disk(RandomFloat2(seed)); }
GK_DEV float3 RandomInCone(uint4& seed, float cos_theta) {  // :282-289
    const float2 u = RandomFloat2(seed);
    float phi = GK_M_TWO_PI * u.x;
    cos_theta = 1.0f + u.y * (cos_theta - 1.f);
    float s, c;
    gk_det_sincosf(phi, &s, &c);
    float r = sqrtf(1.0f - cos_theta * cos_theta);
    return f3(r * c, r * s, cos_theta);
}
GK_DEV float3 RandomInHemiSphere1(uint4& seed) {  // :299-305
    const float2 u = RandomFloat2(seed);
    float phi = GK_M_TWO_PI * u.x;
    float s, c;
    gk_det_sincosf(phi, &s, &c);
    float r = sqrtf(u.y);
    return f3(r * c, r * s, sqrtf(1.0f - u.y));
}
GK_DEV float3 ggx_sample_vndf(float2 alpha, float3 wi_, float2 uv) {  // :310-328
    float3 wi = normalize(f3(wi_.x * alpha.x, wi_.y * alpha.y, wi_.z));
    float b = wi.z;
    if (wi_.z > 0.f) {
        float a = saturatef(fminf(alpha.x, alpha.y));
        float awiz_s = a * wi_.z / (1.0f + length(make_float2(wi_.x, wi_.y)));
        b *= ((1.0f - a * a) / (1.0f + awiz_s * awiz_s));
    }
    float z = __fmaf_rn(1.0f - uv.y, 1.0f + b, -b);
    float phi = GK_M_TWO_PI * uv.x;
    float s, c;
    gk_det_sincosf(phi, &s, &c);
    float r = sqrtf(saturatef(1.0f - z * z));
    float3 o_std = f3(r * c, r * s, z);
    float3 m_std = wi + o_std;
    return normalize(f3(m_std.x * alpha.x, m_std.y * alpha.y, m_std.z));
}
GK_DEV float3 ggxSampling(uint4& RandomSeed, float roughness, float3 normal) {  // :330-340
    float3 tangent, bitangent;
    ONB(normal, tangent, bitangent);
    float a = roughness * roughness;
    float3 wm_ = ggx_sample_vndf(make_float2(a, a), to_local(normal, tangent, bitangent, normal), RandomFloat2(RandomSeed));
    return to_world(wm_, tangent, bitangent, normal);
}
GK_DEV float3 reflect3(float3 i, float3 n) { return i - n * (2.0f * dot(n, i)); }
GK_DEV float3 refract3(float3 i, float3 n, float eta) {
    float ni = dot(n, i);
    float k = 1.0f - eta * eta * (1.0f - ni * ni);
    if (k < 0.0f) return f3(0, 0, 0);
    return i * eta - n * (eta * ni + sqrtf(k));
}

}  // namespace gkd
