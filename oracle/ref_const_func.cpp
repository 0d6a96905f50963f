// Compiles the reference's own sampling / RNG / tonemapping source (assets/shaders/common/Const_Func.slang and PreProcessor.slang) as
// C++ so that tests can pin the oracle's restatements of SURVEY 8a rows a19 and a25 against it. The Slang file is not valid C++ as a
// whole (swizzles, resource types), so oracle/Makefile extracts the line ranges of the functions below with sed into
// oracle/_ref/const_func_gen.inc (a build product next to the other _ref outputs, never committed) and only rewrites the parameter
// qualifiers (`inout T x` / `out T x` -> `T& x`, `public` dropped); the function bodies are the reference's text. This file supplies
// the HLSL vocabulary they use. Floating literals are single precision as in Slang (-fsingle-precision-constant), products and sums
// are not contracted (-ffp-contract=off), dot / mul / normalize are defined the way the oracle defines them (left-to-right, unfused).
// Built only when /root/reference is mounted. TEST INFRASTRUCTURE ONLY.
#include <cmath>
#include <cstdint>
#include <cstring>

#undef M_PI
#undef M_PI_2
#undef M_PI_4
#undef M_1_PI
#undef M_2_PI

namespace ref {
using std::abs;
using std::cos;
using std::fma;
using std::pow;
using std::sin;
using std::sqrt;
typedef uint32_t uint;

struct float2 {
    float x, y;
    float2() : x(0), y(0) {}
    float2(float v) : x(v), y(v) {}
    float2(float a, float b) : x(a), y(b) {}
    float2& operator+=(float2 o) { x += o.x; y += o.y; return *this; }
};
inline float2 operator+(float2 a, float2 b) { return float2(a.x + b.x, a.y + b.y); }
inline float2 operator-(float2 a, float2 b) { return float2(a.x - b.x, a.y - b.y); }
inline float2 operator-(float2 a, float b) { return float2(a.x - b, a.y - b); }
inline float2 operator*(float2 a, float2 b) { return float2(a.x * b.x, a.y * b.y); }
inline float2 operator*(float a, float2 b) { return float2(a * b.x, a * b.y); }
inline float2 operator*(float2 a, float b) { return float2(a.x * b, a.y * b); }
inline float length(float2 a) { return sqrt(a.x * a.x + a.y * a.y); }

struct float3 {
    union {
        struct { float x, y, z; };
        struct { float r, g, b; };
        float2 xy;
    };
    float3() : x(0), y(0), z(0) {}
    float3(float a, float b_, float c) : x(a), y(b_), z(c) {}
    float3(float2 a, float c) : x(a.x), y(a.y), z(c) {}
    float3& operator*=(float s) { x *= s; y *= s; z *= s; return *this; }
};
inline float3 operator+(float3 a, float3 b) { return float3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline float3 operator+(float a, float3 b) { return float3(a + b.x, a + b.y, a + b.z); }
inline float3 operator*(float3 a, float3 b) { return float3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline float3 operator*(float a, float3 b) { return float3(a * b.x, a * b.y, a * b.z); }
inline float3 operator*(float3 a, float b) { return float3(a.x * b, a.y * b, a.z * b); }
inline float3 operator/(float3 a, float b) { return float3(a.x / b, a.y / b, a.z / b); }
inline float3 operator/(float a, float3 b) { return float3(a / b.x, a / b.y, a / b.z); }
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
    };
    float4() : x(0), y(0), z(0), w(0) {}
    float4(float v) : x(v), y(v), z(v), w(v) {}
    float4(float a_, float b_, float c, float d) : x(a_), y(b_), z(c), w(d) {}
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
inline uint2 operator>>(uint2 a, int s) { return uint2{a.x >> s, a.y >> s}; }
inline uint2 operator|(int a, uint2 b) { return uint2{(uint)a | b.x, (uint)a | b.y}; }
struct uint4 {
    union {
        struct { uint x, y, z, w; };
        uint2 xy;
    };
    uint4() : x(0), y(0), z(0), w(0) {}
    uint4(uint a, uint b, uint c, uint d) : x(a), y(b), z(c), w(d) {}
};
inline uint4 operator*(uint4 a, uint s) { return uint4(a.x * s, a.y * s, a.z * s, a.w * s); }
inline uint4 operator+(uint4 a, uint s) { return uint4(a.x + s, a.y + s, a.z + s, a.w + s); }
inline uint4 operator>>(uint4 a, uint s) { return uint4(a.x >> s, a.y >> s, a.z >> s, a.w >> s); }
inline uint4 operator^(uint4 a, uint4 b) { return uint4(a.x ^ b.x, a.y ^ b.y, a.z ^ b.z, a.w ^ b.w); }
inline float asfloat(uint v) { float f; std::memcpy(&f, &v, 4); return f; }
inline float2 asfloat(uint2 v) { return float2(asfloat(v.x), asfloat(v.y)); }

#include "PreProcessor.slang"
#include "const_func_gen.inc"
struct AmbientCube {  // BasicTypes.slang:115-137; the layout itself is pinned by ref_layout_check
    uint PosZ, NegZ, PosY, NegY, PosX, NegX, PosZ_D, NegZ_D, PosY_D, NegY_D, PosX_D, NegX_D, skyVisibility_pznzpyny, skyVisibility_pxnxs0s1;
};
#include "ambient_cube_gen.inc"
}  // namespace ref

using namespace ref;
extern "C" {
float ref_schlick(float cosine, float ior) { return Schlick(cosine, ior); }
void ref_align_with_normal(const float* ray, const float* n, float* out) {
    float3 r = AlignWithNormal(float3(ray[0], ray[1], ray[2]), float3(n[0], n[1], n[2]));
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
void ref_st2084(const float* in, float* out) { float3 r = LinearToST2084UE(float3(in[0], in[1], in[2])); out[0] = r.x; out[1] = r.y; out[2] = r.z; }
void ref_gt_tonemap(const float* in, float* out) { float3 r = GT_Tonemapping(float3(in[0], in[1], in[2])); out[0] = r.x; out[1] = r.y; out[2] = r.z; }
void ref_pcg4d(uint32_t* v) { uint4 s(v[0], v[1], v[2], v[3]); pcg4d(s); v[0] = s.x; v[1] = s.y; v[2] = s.z; v[3] = s.w; }
float ref_uint_to_float(uint32_t x) { return uint_to_float(x); }
float ref_random_float(uint32_t* v) { uint4 s(v[0], v[1], v[2], v[3]); float r = RandomFloat(s); v[0] = s.x; v[1] = s.y; v[2] = s.z; v[3] = s.w; return r; }
void ref_random_float2(uint32_t* v, float* out) {
    uint4 s(v[0], v[1], v[2], v[3]); float2 r = RandomFloat2(s); v[0] = s.x; v[1] = s.y; v[2] = s.z; v[3] = s.w; out[0] = r.x; out[1] = r.y;
}
void ref_random_in_unit_disk(uint32_t* v, float* out) {
    uint4 s(v[0], v[1], v[2], v[3]); float2 r = RandomInUnitDisk(s); v[0] = s.x; v[1] = s.y; v[2] = s.z; v[3] = s.w; out[0] = r.x; out[1] = r.y;
}
void ref_random_in_cone(uint32_t* v, float cos_theta, float* out) {
    uint4 s(v[0], v[1], v[2], v[3]); float3 r = RandomInCone(s, cos_theta); v[0] = s.x; v[1] = s.y; v[2] = s.z; v[3] = s.w; out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
void ref_random_in_hemisphere1(uint32_t* v, float* out) {
    uint4 s(v[0], v[1], v[2], v[3]); float3 r = RandomInHemiSphere1(s); v[0] = s.x; v[1] = s.y; v[2] = s.z; v[3] = s.w; out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
void ref_ggx_sampling(uint32_t* v, float roughness, const float* n, float* out) {
    uint4 s(v[0], v[1], v[2], v[3]); float3 r = ggxSampling(s, roughness, float3(n[0], n[1], n[2]));
    v[0] = s.x; v[1] = s.y; v[2] = s.z; v[3] = s.w; out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
uint32_t ref_pack_rgb10a2(const float* c) { return packRGB10A2(float4(c[0], c[1], c[2], c[3])); }
void ref_unpack_rgb10a2(uint32_t p, float* out) { float4 c = unpackRGB10A2(p); out[0] = c.x; out[1] = c.y; out[2] = c.z; out[3] = c.w; }
uint32_t ref_lerp_packed_color_alt(uint32_t c0, const float* c1, float t) { return LerpPackedColorAlt(c0, float4(c1[0], c1[1], c1[2], c1[3]), t); }
// which: 0 DI, 1 LightCache, 2 Full
void ref_sample_ambient_cube(int which, const uint32_t* cube14, const float* n, float* out) {
    AmbientCube c;
    std::memcpy(&c, cube14, sizeof(c));
    float3 nn(n[0], n[1], n[2]);
    float4 r = which == 0 ? sampleAmbientCubeHL2_DI(c, nn) : which == 1 ? sampleAmbientCubeHL2_LightCache(c, nn) : sampleAmbientCubeHL2_Full(c, nn);
    out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}
}
