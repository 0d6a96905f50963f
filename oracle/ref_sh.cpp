// Compiles the reference's EvaluateSH / SampleIBL (C++ branch of the dual-language header
// assets/shaders/glsl_legacy/common/SampleIBL.glsl:3-50, the same function Shading.slang:146-192 restates for Slang) where the file
// lies under /root/reference, with the few GLSL types it needs shimmed here, and exports it so that tests can pin the oracle's
// SampleIBLRough against the reference's own source. Built by oracle/Makefile (target ref) only when the reference is mounted;
// nothing is copied. Also: ProjectHdrToSh, the host-side SH9 projection of the sky image that produces those coefficients
// (src/Assets/Texture.cpp:180-258 with its M_NEXT_PI, :22), line range copied unmodified into oracle/_ref/project_sh_gen.inc, for pinning
// the host library's restatement (gknextengine_b200/host/gk_scene.cpp ProjectHdrToSh). TEST INFRASTRUCTURE ONLY.
#include <cmath>
#include <cstdint>

#include "../include/gk_detmath.h"

namespace ref {
using uint = uint32_t;
struct vec3 {
    union { float x; float r; };
    union { float y; float g; };
    union { float z; float b; };
    vec3() : x(0), y(0), z(0) {}
    explicit vec3(float v) : x(v), y(v), z(v) {}
    vec3(float a, float b_, float c) : x(a), y(b_), z(c) {}
};
inline vec3 operator*(vec3 a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
struct vec4 {
    float x, y, z, w;
    vec4(vec3 v, float w_) : x(v.x), y(v.y), z(v.z), w(w_) {}
};
// the intrinsics get the pinned kernels of the numeric contract, as in the oracle and the CUDA renderer
inline float cos(float x) { return gk_det_cosf(x); }
inline float sin(float x) { return gk_det_sinf(x); }
struct SphericalHarmonics { float coefficients[3][9]; float padding; };
static SphericalHarmonics HDRSHs[1];
#include "SampleIBL.glsl"
}  // namespace ref

extern "C" void ref_sample_ibl_sh(const float* coefficients27, const float* dir3, float rotate, float* out4) {
    for (int c = 0; c < 3; ++c)
        for (int i = 0; i < 9; ++i) ref::HDRSHs[0].coefficients[c][i] = coefficients27[c * 9 + i];
    ref::vec4 r = ref::SampleIBL(0u, ref::vec3(dir3[0], dir3[1], dir3[2]), rotate, 1.0f);
    out4[0] = r.x; out4[1] = r.y; out4[2] = r.z; out4[3] = r.w;
}

namespace refproj {
struct SphericalHarmonics { float coefficients[3][9]; float padding; };
#include "project_sh_gen.inc"
}  // namespace refproj

extern "C" void ref_project_sh(const float* rgba, int width, int height, float* out27) {
    refproj::SphericalHarmonics sh = refproj::ProjectHdrToSh(rgba, width, height);
    for (int c = 0; c < 3; ++c)
        for (int i = 0; i < 9; ++i) out27[c * 9 + i] = sh.coefficients[c][i];
}
