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
 * The leaf transcendentals - atan2/acos of the sky lookup (Shading.slang:194-199), exp of the bilateral weights
 * (Process.DenoiseJBF.comp.slang), pow of the tonemappers (Const_Func.slang:51-126) - never influence control flow, but two
 * libms round them differently and an fp16 store then flips here and there. They are pinned the same way (Cephes single-precision
 * kernels in plain mul/add, explicit fmaf in exp, IEEE sqrt/div), so that every storage image is bit-identical between the CUDA renderer and its checker:
 * gk_det_atan2f, gk_det_acosf, gk_det_expf, gk_det_logf, gk_det_powf. Errors vs. libm: a few ulp (tests/test_detmath.py); pow is
 * exp(y * log(x)), i.e. relative error ~1e-7 * |y ln x| like a GPU's intrinsic.
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


GK_HD float gk_det_bits2f(unsigned int u) { union { unsigned int u; float f; } c; c.u = u; return c.f; }
GK_HD unsigned int gk_det_f2bits(float f) { union { unsigned int u; float f; } c; c.f = f; return c.u; }

GK_HD float gk_det_atanf(float xx) {
    float sign = 1.0f, x = xx, y;
    if (x < 0.0f) { sign = -1.0f; x = -x; }
    if (x > 2.414213562373095f) { y = 1.5707963267948966f; x = -(1.0f / x); }        /* tan(3 pi / 8) */
    else if (x > 0.4142135623730950f) { y = 0.7853981633974483f; x = (x - 1.0f) / (x + 1.0f); }  /* tan(pi / 8) */
    else y = 0.0f;
    const float z = x * x;
    y = y + ((((8.05374449538e-2f * z - 1.38776856032e-1f) * z + 1.99777106478e-1f) * z - 3.33329491539e-1f) * z * x + x);
    return sign * y;
}

/* atan2(y, x) with the C library's quadrant conventions (NaN in -> NaN out) */
GK_HD float gk_det_atan2f(float y, float x) {
    const float PI = 3.14159265358979323846f, PIO2 = 1.5707963267948966f;
    if (x != x || y != y) return x + y;
    int code = 0;
    if (x < 0.0f) code = 2;
    if (y < 0.0f) code |= 1;
    if (x == 0.0f) {
        if (code & 1) return -PIO2;
        if (y == 0.0f) return 0.0f;
        return PIO2;
    }
    if (y == 0.0f) return (code & 2) ? PI : 0.0f;
    const float w = code == 2 ? PI : (code == 3 ? -PI : 0.0f);
    return w + gk_det_atanf(y / x);
}

GK_HD float gk_det_asinf(float xx) {
    float sign = 1.0f, a = xx, x, z;
    if (a < 0.0f) { sign = -1.0f; a = -a; }
    if (a != a) return xx;  /* NaN in -> the same NaN out, like libm: keeps two compilations of the same formula bit-identical */
    if (a > 1.0f) return gk_det_bits2f(0x7FC00000u);
    if (a < 1.0e-4f) return xx;
    int flag = 0;
    if (a > 0.5f) { z = 0.5f * (1.0f - a); x = sqrtf(z); flag = 1; }
    else { x = a; z = x * x; }
    z = ((((4.2163199048e-2f * z + 2.4181311049e-2f) * z + 4.5470025998e-2f) * z + 7.4953002686e-2f) * z + 1.6666752422e-1f) * z * x + x;
    if (flag) { z = z + z; z = 1.5707963267948966f - z; }
    return sign * z;
}

GK_HD float gk_det_acosf(float x) {
    if (x < -0.5f) return 3.14159265358979323846f - 2.0f * gk_det_asinf(sqrtf(0.5f * (1.0f + x)));
    if (x > 0.5f) return 2.0f * gk_det_asinf(sqrtf(0.5f * (1.0f - x)));
    return 1.5707963267948966f - gk_det_asinf(x);
}

/* exp with gradual underflow. The polynomial and the reduction use explicit fused multiply-adds (fmaf rounds once on both processors), and
 * the scaling by 2^n is two exact power-of-two factors, so the kernel stays short: it runs 42 times per pixel in the bilateral filter. */
GK_HD float gk_det_expf(float x) {
    if (!(x >= -103.9720840454f)) return x != x ? x : 0.0f;  /* below ln(2^-150): 0; NaN in -> the same NaN out */
    if (x > 88.72283905206835f) return gk_det_bits2f(0x7F800000u);
    const float k = floorf(fmaf(x, 1.44269504088896341f, 0.5f));
    float r = fmaf(k, -0.693359375f, x);
    r = fmaf(k, 2.12194440e-4f, r);
    float p = 1.9875691500e-4f;
    p = fmaf(p, r, 1.3981999507e-3f);
    p = fmaf(p, r, 8.3334519073e-3f);
    p = fmaf(p, r, 4.1665795894e-2f);
    p = fmaf(p, r, 1.6666665459e-1f);
    p = fmaf(p, r, 5.0000001201e-1f);
    p = fmaf(p, r * r, r) + 1.0f;
    const int n = (int)k;            /* -150 .. 128 */
    const int n1 = n >> 1, n2 = n - n1; /* both factors are normal numbers; only the last product can round (into the subnormals) */
    return p * gk_det_bits2f((unsigned int)(n1 + 127) << 23) * gk_det_bits2f((unsigned int)(n2 + 127) << 23);
}

/* natural logarithm of a positive finite number; 0 -> -inf, negative / NaN -> NaN, +inf -> +inf */
GK_HD float gk_det_logf(float xx) {
    if (xx != xx) return xx;
    if (xx < 0.0f) return gk_det_bits2f(0x7FC00000u);
    if (xx == 0.0f) return gk_det_bits2f(0xFF800000u);
    unsigned int u = gk_det_f2bits(xx);
    if (u == 0x7F800000u) return xx;
    int e = 0;
    if (u < 0x00800000u) { xx = xx * 33554432.0f; e = -25; u = gk_det_f2bits(xx); }  /* subnormal: scale by 2^25 */
    e += (int)(u >> 23) - 126;                                                     /* xx = m * 2^e with m in [0.5, 1) */
    float x = gk_det_bits2f((u & 0x007FFFFFu) | 0x3F000000u);
    if (x < 0.707106781186547524f) { e -= 1; x = x + x - 1.0f; }
    else x = x - 1.0f;
    float z = x * x;
    float y = ((((((((7.0376836292e-2f * x - 1.1514610310e-1f) * x + 1.1676998740e-1f) * x - 1.2420140846e-1f) * x + 1.4249322787e-1f) * x - 1.6668057665e-1f) * x +
                 2.0000714765e-1f) * x - 2.4999993993e-1f) * x + 3.3333331174e-1f) * x * z;
    const float fe = (float)e;
    y = y + -2.12194440e-4f * fe;
    y = y + -0.5f * z;
    z = x + y;
    z = z + 0.693359375f * fe;
    return z;
}

/* pow for the tonemappers: x >= 0 expected. pow(0, y > 0) = 0, pow(x, 0) = 1, pow(negative, y) = NaN, pow(+inf, y > 0) = +inf */
GK_HD float gk_det_powf(float x, float y) {
    if (y == 0.0f) return 1.0f;
    if (x != x) return x;
    if (y != y) return y;
    if (x < 0.0f) return gk_det_bits2f(0x7FC00000u);
    if (x == 0.0f) return y > 0.0f ? 0.0f : gk_det_bits2f(0x7F800000u);
    return gk_det_expf(y * gk_det_logf(x));
}

#endif /* GK_DETMATH_H */
