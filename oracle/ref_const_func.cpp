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
}
