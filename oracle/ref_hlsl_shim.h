// HLSL / Slang vocabulary for compiling extracts of the reference's shader sources as C++ (oracle/ref_const_func.cpp,
// oracle/ref_general_func.cpp). Only what those extracts use; dot / mul / mad / normalize are defined the way the oracle defines them
// (left-to-right, unfused) - the pins check that the oracle restates the reference's text, not how a GPU rounds it.
// TEST INFRASTRUCTURE ONLY, used only when /root/reference is mounted.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <utility>

#undef M_PI
#undef M_PI_2
#undef M_PI_4
#undef M_1_PI
#undef M_2_PI

#ifdef REF_DET_MATH
#include "../include/gk_detmath.h"
#endif

namespace ref {
using std::abs;
using std::fma;
#ifdef REF_DET_MATH
// sin / cos / atan2 / acos / exp / pow are hardware intrinsics in the shaders, with driver-defined rounding. The pins give them the pinned
// kernels of the numeric contract (include/gk_detmath.h) that the oracle and the CUDA renderer use, so that whole passes compare bit for bit.
inline float sin(float x) { return gk_det_sinf(x); }
inline float cos(float x) { return gk_det_cosf(x); }
inline float atan2(float y, float x) { return gk_det_atan2f(y, x); }
inline float acos(float x) { return gk_det_acosf(x); }
inline float exp(float x) { return gk_det_expf(x); }
inline float pow(float x, float y) { return gk_det_powf(x, y); }
#else
using std::cos;
using std::exp;
using std::pow;
using std::sin;
#endif
using std::sqrt;
typedef uint32_t uint;

struct float2 {
    float x, y;
    float2() : x(0), y(0) {}
    float2(float v) : x(v), y(v) {}
    float2(float a, float b) : x(a), y(b) {}
    template <class V, class = decltype(std::declval<V>().xy)>
    explicit float2(const V& v) : x((float)v.x), y((float)v.y) {}  // float2(int2)
    float2& operator+=(float2 o) { x += o.x; y += o.y; return *this; }
};
inline float2 operator+(float2 a, float2 b) { return float2(a.x + b.x, a.y + b.y); }
inline float2 operator-(float2 a, float2 b) { return float2(a.x - b.x, a.y - b.y); }
inline float2 operator-(float2 a, float b) { return float2(a.x - b, a.y - b); }
inline float2 operator*(float2 a, float2 b) { return float2(a.x * b.x, a.y * b.y); }
inline float2 operator*(float a, float2 b) { return float2(a * b.x, a * b.y); }
inline float2 operator*(float2 a, float b) { return float2(a.x * b, a.y * b); }
inline float length(float2 a) { return sqrt(a.x * a.x + a.y * a.y); }

struct float3;
struct float3rgb { float x, y, z; inline operator float3() const; inline float3rgb& operator=(const float3& v); };  // `.rgb` of a float3
struct float3 {
    union {
        struct { float x, y, z; };
        struct { float r, g, b; };
        float2 xy;
        float3rgb rgb;
        float3rgb xyz;
    };
    float3() : x(0), y(0), z(0) {}
    float3(float a, float b_, float c) : x(a), y(b_), z(c) {}
    float3(float2 a, float c) : x(a.x), y(a.y), z(c) {}
    explicit float3(float v) : x(v), y(v), z(v) {}
    float3& operator*=(float s) { x *= s; y *= s; z *= s; return *this; }
};
inline float3rgb::operator float3() const { return float3(x, y, z); }
inline float3rgb& float3rgb::operator=(const float3& v) { x = v.x; y = v.y; z = v.z; return *this; }
inline float3 operator+(float3 a, float3 b) { return float3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline float3 operator+(float a, float3 b) { return float3(a + b.x, a + b.y, a + b.z); }
inline float3 operator*(float3 a, float3 b) { return float3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline float3 operator*(float a, float3 b) { return float3(a * b.x, a * b.y, a * b.z); }
inline float3 operator*(float3 a, float b) { return float3(a.x * b, a.y * b, a.z * b); }
inline float3 operator/(float3 a, float b) { return float3(a.x / b, a.y / b, a.z / b); }
inline float3 operator/(float a, float3 b) { return float3(a / b.x, a / b.y, a / b.z); }
inline float3 operator-(float3 a, float3 b) { return float3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline float3 operator-(float3 a, float b) { return float3(a.x - b, a.y - b, a.z - b); }
inline float3 cross(float3 a, float3 b) { return float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
inline float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float3 normalize(float3 a) { float inv = 1.0f / sqrt(dot(a, a)); return a * inv; }
inline float3 pow(float3 a, float3 b) { return float3(pow(a.x, b.x), pow(a.y, b.y), pow(a.z, b.z)); }
inline float min(float a, float b) { return fminf(a, b); }
inline float clamp(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }

struct float4 {
    union {
        struct { float x, y, z, w; };
        struct { float r, g, b, a; };
        float3 xyz;
        float3 rgb;
        float2 xy;
        float2 rg;
    };
    float4() : x(0), y(0), z(0), w(0) {}
    float4(float v) : x(v), y(v), z(v), w(v) {}
    float4(float a_, float b_, float c, float d) : x(a_), y(b_), z(c), w(d) {}
    float4(float3 v, float d) : x(v.x), y(v.y), z(v.z), w(d) {}
    float4(float2 v, float c, float d) : x(v.x), y(v.y), z(c), w(d) {}
    float4& operator+=(float4 o) { x += o.x; y += o.y; z += o.z; w += o.w; return *this; }
    float4& operator*=(float s) { x *= s; y *= s; z *= s; w *= s; return *this; }
};
inline float4 operator+(float4 a, float4 b) { return float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
inline float4 operator-(float4 a, float4 b) { return float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
inline float4 operator*(float4 a, float s) { return float4(a.x * s, a.y * s, a.z * s, a.w * s); }
inline float4 operator*(float s, float4 a) { return float4(s * a.x, s * a.y, s * a.z, s * a.w); }
inline float4 operator/(float4 a, float s) { return float4(a.x / s, a.y / s, a.z / s, a.w / s); }
inline float4 clamp(float4 v, float4 lo, float4 hi) { return float4(clamp(v.x, lo.x, hi.x), clamp(v.y, lo.y, hi.y), clamp(v.z, lo.z, hi.z), clamp(v.w, lo.w, hi.w)); }
inline float4 lerp(float4 a, float4 b, float t) { return a + (b - a) * t; }
inline float max(float a, float b) { return fmaxf(a, b); }

inline float4 operator*(float4 a, float4 b) { return float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
inline float2 mad(float2 a, float2 b, float2 c) { return a * b + c; }
inline float3 mad(float3 a, float3 b, float3 c) { return a * b + c; }
inline float4 mad(float4 a, float4 b, float4 c) { return a * b + c; }
struct float4x4 { float m[4][4]; };  // glm column-major storage m[col][row]; mul(M, v) = M * v as the shaders see the uploaded matrices
inline float4 mul(const float4x4& M, float4 v) {
    return float4(M.m[0][0] * v.x + M.m[1][0] * v.y + M.m[2][0] * v.z + M.m[3][0] * v.w, M.m[0][1] * v.x + M.m[1][1] * v.y + M.m[2][1] * v.z + M.m[3][1] * v.w,
                  M.m[0][2] * v.x + M.m[1][2] * v.y + M.m[2][2] * v.z + M.m[3][2] * v.w, M.m[0][3] * v.x + M.m[1][3] * v.y + M.m[2][3] * v.z + M.m[3][3] * v.w);
}
// IEEE binary16 storage (exact widening)
struct half {
    uint16_t bits;
    operator float() const {
        const uint32_t s = (bits >> 15) & 1u, e = (bits >> 10) & 31u, m = bits & 1023u;
        float v;
        if (e == 0) v = std::ldexp((float)m, -24);
        else if (e == 31) v = m ? NAN : INFINITY;
        else v = std::ldexp((float)(m | 1024u), (int)e - 25);
        return s ? -v : v;
    }
};
struct half3 { half x, y, z; operator float3() const { return float3((float)x, (float)y, (float)z); } };
struct half4 {
    union {
        struct { half x, y, z, w; };
        half3 xyz;
    };
    operator float4() const { return float4((float)x, (float)y, (float)z, (float)w); }
};
inline uint16_t asuint16(half h) { return h.bits; }

struct float3x3 {
    float3 row[3];
    float3x3(float3 a, float3 b, float3 c) { row[0] = a; row[1] = b; row[2] = c; }
};
inline float3 mul(float3x3 M, float3 v) { return float3(dot(M.row[0], v), dot(M.row[1], v), dot(M.row[2], v)); }
inline float3 mul(float3 v, float3x3 M) {
    return float3(v.x * M.row[0].x + v.y * M.row[1].x + v.z * M.row[2].x, v.x * M.row[0].y + v.y * M.row[1].y + v.z * M.row[2].y,
                  v.x * M.row[0].z + v.y * M.row[1].z + v.z * M.row[2].z);
}

struct uint2 { uint x, y; };
inline uint2 make_uint2(uint a, uint b) { return uint2{a, b}; }
inline uint2 operator>>(uint2 a, int s) { return uint2{a.x >> s, a.y >> s}; }
inline uint2 operator|(int a, uint2 b) { return uint2{(uint)a | b.x, (uint)a | b.y}; }
struct uint4 {
    union {
        struct { uint x, y, z, w; };
        uint2 xy;
    };
    uint4() : x(0), y(0), z(0), w(0) {}
    uint4(uint a, uint b, uint c, uint d) : x(a), y(b), z(c), w(d) {}
    explicit uint4(uint v) : x(v), y(v), z(v), w(v) {}
};
inline uint4 operator*(uint4 a, uint s) { return uint4(a.x * s, a.y * s, a.z * s, a.w * s); }
inline uint4 operator+(uint4 a, uint s) { return uint4(a.x + s, a.y + s, a.z + s, a.w + s); }
inline uint4 operator>>(uint4 a, uint s) { return uint4(a.x >> s, a.y >> s, a.z >> s, a.w >> s); }
inline uint4 operator^(uint4 a, uint4 b) { return uint4(a.x ^ b.x, a.y ^ b.y, a.z ^ b.z, a.w ^ b.w); }
struct bool4 { bool x, y, z, w; };
inline bool any(bool4 b) { return b.x || b.y || b.z || b.w; }
inline uint4 splat4(uint v) { return uint4(v, v, v, v); }
inline bool4 operator!=(uint4 a, uint4 b) { return bool4{a.x != b.x, a.y != b.y, a.z != b.z, a.w != b.w}; }
inline bool4 operator==(uint4 a, uint4 b) { return bool4{a.x == b.x, a.y == b.y, a.z == b.z, a.w == b.w}; }
inline float asfloat(uint v) { float f; std::memcpy(&f, &v, 4); return f; }
inline float2 asfloat(uint2 v) { return float2(asfloat(v.x), asfloat(v.y)); }

}  // namespace ref
