// Compiles the reference's own sampling / RNG / tonemapping source (assets/shaders/common/Const_Func.slang and PreProcessor.slang) as
// C++ so that tests can pin the oracle's restatements of SURVEY 8a rows a19 and a25 against it. The Slang file is not valid C++ as a
// whole (swizzles, resource types), so oracle/Makefile extracts the line ranges of the functions below with sed into
// oracle/_ref/const_func_gen.inc (a make intermediate: deleted again as soon as the library is linked, never committed) and only rewrites the parameter
// qualifiers (`inout T x` / `out T x` -> `T& x`, `public` dropped); the function bodies are the reference's text. This file supplies
// the HLSL vocabulary they use. Floating literals are single precision as in Slang (-fsingle-precision-constant), products and sums
// are not contracted (-ffp-contract=off), dot / mul / normalize are defined the way the oracle defines them (left-to-right, unfused).
// Built only when /root/reference is mounted. TEST INFRASTRUCTURE ONLY.
#define REF_DET_MATH
#include "ref_hlsl_shim.h"

namespace ref {
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
