/*
 * gk_detmath.h — the numeric contract shared by the CUDA renderer and its CPU checker.
 *
 * The reference shaders call sin/cos on random angles to build ray directions
 * (assets/shaders/common/Const_Func.slang:282-328); the result feeds BVH traversal and every later
 * random decision of the path, so a 1-ulp difference between two libm implementations makes two
 * otherwise identical renderers diverge. Slang/SPIR-V leaves the precision of these intrinsics to the
 * driver, so there is no "reference" bit pattern to match; this header pins one: a Cephes-style
 * Cody-Waite reduction + minimax polynomial evaluated in plain fp32 mul/add (no FMA contraction:
 * nvcc -fmad=false, g++ -ffp-contract=off). sqrt and division are IEEE-exact on both sides.
 * Max error vs. libm is < 2 ulp on [0, 2*pi] (tests/test_detmath.py).
 *
 * Leaf-only transcendentals (atan2/acos for the sky lookup, pow/exp in the tonemapper and the
 * bilateral weights) stay on libm / CUDA math: they never influence control flow.
 */
#ifndef GK_DETMATH_H
#define GK_DETMATH_H

#ifdef __CUDACC__
#define GK_HD __host__ __device__ __forceinline__
#else
#define GK_HD inline
#endif

GK_HD void gk_det_sincosf(float xx, float* s, float* c) {
    const float FOPI = 1.27323954473516f; /* 4/pi */
    const float DP1 = 0.78515625f, DP2 = 2.4187564849853515625e-4f, DP3 = 3.77489497744594108e-8f;
    float ssign = 1.0f, csign = 1.0f;
    float x = xx;
    if (x < 0.0f) { ssign = -1.0f; x = -x; }
    int j = (int)(FOPI * x);
    float y = (float)j;
    if (j & 1) { j += 1; y += 1.0f; }
    j &= 7;
    if (j > 3) { ssign = -ssign; csign = -csign; j -= 4; }
    if (j > 1) csign = -csign;
    x = ((x - y * DP1) - y * DP2) - y * DP3;
    float z = x * x;
    float ps = ((-1.9515295891e-4f * z + 8.3321608736e-3f) * z - 1.6666654611e-1f) * z * x + x;
    float pc = ((2.443315711809948e-5f * z - 1.388731625493765e-3f) * z + 4.166664568298827e-2f) * z * z - 0.5f * z + 1.0f;
    if (j == 1 || j == 2) { *s = ssign * pc; *c = csign * ps; }
    else { *s = ssign * ps; *c = csign * pc; }
}

GK_HD float gk_det_sinf(float x) { float s, c; gk_det_sincosf(x, &s, &c); return s; }
GK_HD float gk_det_cosf(float x) { float s, c; gk_det_sincosf(x, &s, &c); return c; }

#endif /* GK_DETMATH_H */
