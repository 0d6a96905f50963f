// gk_post.cu — the two stencil passes after the rt pass, as shared-memory-tiled kernels.
//   k_reproject  = assets/shaders/Process.ReProject.comp.slang:71-223   ("reproject pass", PathTracingRenderer.cpp:126-136)
//   k_denoise    = assets/shaders/Process.DenoiseJBF.comp.slang:75-182  ("compose pass",   PathTracingRenderer.cpp:138-157)
// Storage images are fp16x4 / R32_UINT / RG32F exactly as in the reference (VulkanBaseRenderer.cpp:439-457); out-of-bounds
// loads return 0 and out-of-bounds stores are dropped, like Vulkan robust image access.
// The "copy pass" (ACCUM -> PREV x3, PathTracingRenderer.cpp:160-182) and the object-id copy (VulkanBaseRenderer.cpp:1136-1152)
// are pointer swaps done by the caller; ReProject therefore writes ACCUM.a = PREV.a where the reference leaves alpha untouched.
#include <algorithm>

#include "gk_context.h"
#include "gk_tma.cuh"

using namespace gkd;

namespace gkcu {

namespace {

constexpr int RBX = 32, RBY = 8;            // reproject block
constexpr int RH = 2;                       // 5x5 spatial filter halo
constexpr int RTW = RBX + 2 * RH, RTH = RBY + 2 * RH;
constexpr int RMH = 4;                      // motion-vector taps at 2 * offset
constexpr int RMW = RBX + 2 * RMH, RMHH = RBY + 2 * RMH;
// TMA: the first column of a box must start on a 16-byte boundary in global memory (measured on sm_100a: a box starting at
// column -2 of a 4-byte image, or column -5 of an 8-byte image, raises "illegal instruction"; scripts/tma_test3.cu), so the
// object-id tile is loaded with a 4-column halo and the denoise tile with a 6-column halo.
constexpr int RIH = 4, RIW = RBX + 2 * RIH;

struct ReprojArgs {
    int w, h;
    const uint2 *single_diffuse, *single_specular, *albedo, *normal;
    const uint2 *prev_diffuse, *prev_specular, *prev_albedo;
    uint2 *accum_diffuse, *accum_specular, *accum_albedo;
    const uint32_t *objectid0, *objectid1;
    const float2* motion;
    const uint16_t* moment;
    uint32_t progressive, fast_reproject;
    int temporal_frames;
    // strip-parallel post (multi-GPU): only rows [y_begin, y_end) are processed; rows [own_y0, own_y1) are also stored into the ACCUM_*
    // images of n_mirror other ranks over peer memory, so that every rank holds the whole history for the next frame's reprojection
    int y_begin, y_end, own_y0, own_y1, n_mirror;
    uint2 *mirror_diffuse[GKCU_MAX_PEERS - 1], *mirror_specular[GKCU_MAX_PEERS - 1], *mirror_albedo[GKCU_MAX_PEERS - 1];
};

GK_DEV void store_accum(const ReprojArgs& A, size_t idx, int y, float4 d, float4 s, float4 a) {
    store_h4(A.accum_diffuse, idx, d);
    store_h4(A.accum_specular, idx, s);
    store_h4(A.accum_albedo, idx, a);
    if (A.n_mirror && y >= A.own_y0 && y < A.own_y1) {
        const uint2 dd = A.accum_diffuse[idx], ss = A.accum_specular[idx], aa = A.accum_albedo[idx];  // the fp16 bits just stored
        for (int r = 0; r < A.n_mirror; ++r) {
            A.mirror_diffuse[r][idx] = dd;
            A.mirror_specular[r][idx] = ss;
            A.mirror_albedo[r][idx] = aa;
        }
    }
}

GK_DEV bool inb(int x, int y, int w, int h) { return x >= 0 && y >= 0 && x < w && y < h; }
GK_DEV float4 ld4(const uint2* img, int x, int y, int w, int h) { return inb(x, y, w, h) ? load_h4(img, (size_t)y * w + x) : f4(0, 0, 0, 0); }
GK_DEV uint32_t ldu(const uint32_t* img, int x, int y, int w, int h) { return inb(x, y, w, h) ? img[(size_t)y * w + x] : 0u; }

GK_DEV float calculateWeight(float centerDist, bool sameObject, bool isCenter, float3 Normal, float3 CenterNormal) {  // :32-47
    if (isCenter) return 0.4f;
    if (!sameObject) return 0.0f;
    float normalDot = clampf(dot(Normal, CenterNormal), 0.0f, 1.0f);
    const float normalThreshold = 0.98f;
    if (normalDot < normalThreshold) return 0.0f;
    const float normalWeight = (normalDot - normalThreshold) / (1.0f - normalThreshold);
    return normalWeight * 2.0f / (centerDist * 1.5f + 4.0f);
}

GK_DEV float3 FilterHistoryColor(const uint2* prev, int w, int h, float2 subpixel, int px, int py, uint32_t cur, uint32_t p0, uint32_t p1, uint32_t p2,
                                 uint32_t p3, float4 currFrameSample) {  // :49-68
    float3 c = xyz(currFrameSample);
    float3 h0 = cur == p0 ? xyz(ld4(prev, px, py, w, h)) : c;
    float3 h1 = cur == p1 ? xyz(ld4(prev, px + 1, py, w, h)) : c;
    float3 h2 = cur == p2 ? xyz(ld4(prev, px, py + 1, w, h)) : c;
    float3 h3 = cur == p3 ? xyz(ld4(prev, px + 1, py + 1, w, h)) : c;
    float3 history = lerp3(lerp3(h0, h1, subpixel.x), lerp3(h2, h3, subpixel.x), subpixel.y);
    return f3(clampf(history.x, 0.0f, 1600.0f), clampf(history.y, 0.0f, 1600.0f), clampf(history.z, 0.0f, 1600.0f));
}

__global__ void __launch_bounds__(RBX* RBY) k_reproject(ReprojArgs A) {
    __shared__ uint2 s_diff[RTH][RTW];
    __shared__ uint2 s_spec[RTH][RTW];
    __shared__ uint2 s_norm[RTH][RTW];
    __shared__ uint32_t s_id[RTH][RTW];
    __shared__ float s_mlen[RMHH][RMW];
    const int w = A.w, h = A.h;
    const int bx0 = blockIdx.x * RBX, by0 = A.y_begin + blockIdx.y * RBY;
    const int tid = threadIdx.y * RBX + threadIdx.x;
    const int x = bx0 + threadIdx.x, y = by0 + threadIdx.y;
    const float k512 = clampf(1.0f / 512.0f, 0.0f, 1.0f);
    if (A.progressive) {  // :95-101
        if (x < w && y < A.y_end) {
            size_t i = (size_t)y * w + x;
            store_accum(A, i, y, lerp4(load_h4(A.prev_diffuse, i), load_h4(A.single_diffuse, i), k512),
                        lerp4(load_h4(A.prev_specular, i), load_h4(A.single_specular, i), k512), lerp4(load_h4(A.prev_albedo, i), load_h4(A.albedo, i), k512));
        }
        return;
    }
    // stage the tile + halo (raw fp16 bits; zero outside the image)
    for (int i = tid; i < RTW * RTH; i += RBX * RBY) {
        int ty = i / RTW, tx = i % RTW;
        int gx = bx0 + tx - RH, gy = by0 + ty - RH;
        bool in = inb(gx, gy, w, h);
        size_t g = (size_t)gy * w + gx;
        s_diff[ty][tx] = in ? A.single_diffuse[g] : make_uint2(0, 0);
        s_spec[ty][tx] = in ? A.single_specular[g] : make_uint2(0, 0);
        s_norm[ty][tx] = in ? A.normal[g] : make_uint2(0, 0);
        s_id[ty][tx] = in ? A.objectid0[g] : 0u;
    }
    for (int i = tid; i < RMW * RMHH; i += RBX * RBY) {
        int ty = i / RMW, tx = i % RMW;
        int gx = bx0 + tx - RMH, gy = by0 + ty - RMH;
        float2 m = inb(gx, gy, w, h) ? A.motion[(size_t)gy * w + gx] : make_float2(0, 0);
        s_mlen[ty][tx] = length(m);
    }
    __syncthreads();
    if (x >= w || y >= A.y_end) return;
    const size_t idx = (size_t)y * w + x;
    const int lx = threadIdx.x + RH, ly = threadIdx.y + RH;
    auto h4 = [](uint2 v) {
        __half2 a = *reinterpret_cast<__half2*>(&v.x), b = *reinterpret_cast<__half2*>(&v.y);
        float2 fa = __half22float2(a), fb = __half22float2(b);
        return f4(fa.x, fa.y, fb.x, fb.y);
    };
    float4 currDiffuseSample = h4(s_diff[ly][lx]), currSpecularSample = h4(s_spec[ly][lx]), currAlbedoSample = load_h4(A.albedo, idx);
    bool useHistory = true;
    float2 motion = A.motion[idx];
    int pvx = (int)floorf((float)x + motion.x), pvy = (int)floorf((float)y + motion.y);
    // all(previpos < isize) mixes int and uint: negative coordinates fail the test (C conversion rules)
    const bool inside = ((uint32_t)pvx < (uint32_t)w) && ((uint32_t)pvy < (uint32_t)h) && pvx >= -1 && pvy >= -1;
    uint32_t cur = s_id[ly][lx];
    uint32_t prev0 = ldu(A.objectid1, pvx, pvy, w, h);
    const bool miss = cur == 65535u;
    uint32_t moment = A.moment[idx];
    if (miss || !inside || moment > 0) useHistory = false;
    if (useHistory) {
        const int SAMPLE_RANGE = A.fast_reproject ? 0 : 2;
        const float3 CenterNormal = xyz(h4(s_norm[ly][lx]));
        float4 spatialDiffuseSample = f4(0, 0, 0, 0), spatialSpecularSample = f4(0, 0, 0, 0);
        float totalWeight = 0.0f, motionInRange = 0;
        for (int yy = -SAMPLE_RANGE; yy <= SAMPLE_RANGE; yy++) {
            for (int xx = -SAMPLE_RANGE; xx <= SAMPLE_RANGE; xx++) {
                const uint32_t primitiveIdx = s_id[ly + yy][lx + xx];
                const float3 worldNormal = xyz(h4(s_norm[ly + yy][lx + xx]));
                const bool isCenter = (xx == 0 && yy == 0);
                const bool sameObject = (primitiveIdx == cur);
                const float centerDist = length(make_float2((float)xx, (float)yy));
                const float weight = calculateWeight(centerDist, sameObject, isCenter, worldNormal, CenterNormal);
                spatialDiffuseSample = spatialDiffuseSample + h4(s_diff[ly + yy][lx + xx]) * weight;
                spatialSpecularSample = spatialSpecularSample + h4(s_spec[ly + yy][lx + xx]) * weight;
                totalWeight += weight;
                motionInRange += s_mlen[threadIdx.y + RMH + yy * 2][threadIdx.x + RMH + xx * 2];
            }
        }
        spatialDiffuseSample = spatialDiffuseSample / totalWeight;
        spatialSpecularSample = spatialSpecularSample / totalWeight;
        uint32_t prev1 = ldu(A.objectid1, pvx + 1, pvy, w, h), prev2 = ldu(A.objectid1, pvx, pvy + 1, w, h), prev3 = ldu(A.objectid1, pvx + 1, pvy + 1, w, h);
        if (fabsf(motionInRange) < 0.1f) prev0 = prev1 = prev2 = prev3 = cur;
        float2 pos = make_float2((float)x + motion.x, (float)y + motion.y);
        float2 subpixel = make_float2(pos.x - floorf(pos.x), pos.y - floorf(pos.y));
        float3 historyDiffuse = FilterHistoryColor(A.prev_diffuse, w, h, subpixel, pvx, pvy, cur, prev0, prev1, prev2, prev3, spatialDiffuseSample);
        float3 historySpecular = FilterHistoryColor(A.prev_specular, w, h, subpixel, pvx, pvy, cur, prev0, prev1, prev2, prev3, spatialSpecularSample);
        float3 historyAlbedo = FilterHistoryColor(A.prev_albedo, w, h, subpixel, pvx, pvy, cur, prev0, prev1, prev2, prev3, currAlbedoSample);
        float currKeep = clampf(1.0f / (float)max(1, A.temporal_frames), 0.0f, 1.0f);
        float3 d = lerp3(historyDiffuse, xyz(currDiffuseSample), currKeep);
        float3 s = lerp3(historySpecular, xyz(currSpecularSample), currKeep);
        float3 a = lerp3(historyAlbedo, xyz(currAlbedoSample), currKeep);
        // ".rgb =" stores keep the texel's alpha, which after the copy pass equals PREV's alpha
        store_accum(A, idx, y, f4(d.x, d.y, d.z, load_h4(A.prev_diffuse, idx).w), f4(s.x, s.y, s.z, load_h4(A.prev_specular, idx).w),
                    f4(a.x, a.y, a.z, load_h4(A.prev_albedo, idx).w));
    } else {
        store_accum(A, idx, y, currDiffuseSample, currSpecularSample, currAlbedoSample);
    }
}


// ------------------------------------------------------------------------------------------------ ReProject, TMA-staged
// Same arithmetic as k_reproject, different staging: one elected thread issues five cp.async.bulk.tensor copies (tile + halo of
// SINGLE_DIFFUSE, SINGLE_SPECULAR, NORMAL, OBJECTID_0 and the wider motion-vector tile) onto one mbarrier; the hardware
// zero-fills everything outside the image. The raw fp16 tiles are converted to float once (every texel is read by up to 25
// threads), and the tap loops are unrolled at compile time (RANGE = 2, or 0 for FastReproject).
struct ReprojMaps { CUtensorMap diff, spec, norm, ids, motion; };

template <int RANGE>
__global__ void __launch_bounds__(RBX* RBY) k_reproject_tma(const __grid_constant__ ReprojMaps M, ReprojArgs A) {
    __shared__ alignas(128) uint2 r_diff[RTH][RTW];
    __shared__ alignas(128) uint2 r_spec[RTH][RTW];
    __shared__ alignas(128) uint2 r_norm[RTH][RTW];
    __shared__ alignas(128) uint32_t s_id[RTH][RIW];
    __shared__ alignas(128) float2 r_motion[RMHH][RMW];
    __shared__ float4 s_diff[RTH][RTW];
    __shared__ float4 s_spec[RTH][RTW];
    __shared__ float4 s_norm[RTH][RTW];
    __shared__ float s_mlen[RMHH][RMW];
    __shared__ uint64_t bar;
    const int w = A.w, h = A.h;
    const int bx0 = blockIdx.x * RBX, by0 = A.y_begin + blockIdx.y * RBY;
    const int tid = threadIdx.y * RBX + threadIdx.x;
    const int x = bx0 + threadIdx.x, y = by0 + threadIdx.y;
    if (tid == 0) gktma::mbar_init(&bar, 1);
    __syncthreads();
    if (tid == 0) {
        gktma::mbar_expect_tx(&bar, (uint32_t)(sizeof(r_diff) + sizeof(r_spec) + sizeof(r_norm) + sizeof(s_id) + sizeof(r_motion)));
        gktma::tma_load_2d(r_diff, &M.diff, bx0 - RH, by0 - RH, &bar);
        gktma::tma_load_2d(r_spec, &M.spec, bx0 - RH, by0 - RH, &bar);
        gktma::tma_load_2d(r_norm, &M.norm, bx0 - RH, by0 - RH, &bar);
        gktma::tma_load_2d(s_id, &M.ids, bx0 - RIH, by0 - RH, &bar);
        gktma::tma_load_2d(r_motion, &M.motion, bx0 - RMH, by0 - RMH, &bar);
    }
    gktma::mbar_wait(&bar, 0);
    auto h4 = [](uint2 v) {
        __half2 a = *reinterpret_cast<__half2*>(&v.x), b = *reinterpret_cast<__half2*>(&v.y);
        float2 fa = __half22float2(a), fb = __half22float2(b);
        return f4(fa.x, fa.y, fb.x, fb.y);
    };
    for (int i = tid; i < RTW * RTH; i += RBX * RBY) {
        int ty = i / RTW, tx = i % RTW;
        s_diff[ty][tx] = h4(r_diff[ty][tx]);
        s_spec[ty][tx] = h4(r_spec[ty][tx]);
        s_norm[ty][tx] = h4(r_norm[ty][tx]);
    }
    for (int i = tid; i < RMW * RMHH; i += RBX * RBY) {
        int ty = i / RMW, tx = i % RMW;
        s_mlen[ty][tx] = length(r_motion[ty][tx]);
    }
    __syncthreads();
    if (x >= w || y >= A.y_end) return;
    const size_t idx = (size_t)y * w + x;
    const int lx = threadIdx.x + RH, ly = threadIdx.y + RH;
    float4 currDiffuseSample = s_diff[ly][lx], currSpecularSample = s_spec[ly][lx], currAlbedoSample = load_h4(A.albedo, idx);
    bool useHistory = true;
    float2 motion = r_motion[threadIdx.y + RMH][threadIdx.x + RMH];
    int pvx = (int)floorf((float)x + motion.x), pvy = (int)floorf((float)y + motion.y);
    const bool inside = ((uint32_t)pvx < (uint32_t)w) && ((uint32_t)pvy < (uint32_t)h) && pvx >= -1 && pvy >= -1;
    uint32_t cur = s_id[ly][lx + (RIH - RH)];
    uint32_t prev0 = ldu(A.objectid1, pvx, pvy, w, h);
    const bool miss = cur == 65535u;
    uint32_t moment = A.moment[idx];
    if (miss || !inside || moment > 0) useHistory = false;
    if (useHistory) {
        const float3 CenterNormal = xyz(s_norm[ly][lx]);
        float4 spatialDiffuseSample = f4(0, 0, 0, 0), spatialSpecularSample = f4(0, 0, 0, 0);
        float totalWeight = 0.0f, motionInRange = 0;
#pragma unroll
        for (int yy = -RANGE; yy <= RANGE; yy++) {
#pragma unroll
            for (int xx = -RANGE; xx <= RANGE; xx++) {
                const uint32_t primitiveIdx = s_id[ly + yy][lx + xx + (RIH - RH)];
                const float3 worldNormal = xyz(s_norm[ly + yy][lx + xx]);
                const bool isCenter = (xx == 0 && yy == 0);
                const bool sameObject = (primitiveIdx == cur);
                const float centerDist = sqrtf((float)(xx * xx + yy * yy));  // length(float2(x, y)), folded at compile time
                const float weight = calculateWeight(centerDist, sameObject, isCenter, worldNormal, CenterNormal);
                spatialDiffuseSample = spatialDiffuseSample + s_diff[ly + yy][lx + xx] * weight;
                spatialSpecularSample = spatialSpecularSample + s_spec[ly + yy][lx + xx] * weight;
                totalWeight += weight;
                motionInRange += s_mlen[threadIdx.y + RMH + yy * 2][threadIdx.x + RMH + xx * 2];
            }
        }
        spatialDiffuseSample = spatialDiffuseSample / totalWeight;
        spatialSpecularSample = spatialSpecularSample / totalWeight;
        uint32_t prev1 = ldu(A.objectid1, pvx + 1, pvy, w, h), prev2 = ldu(A.objectid1, pvx, pvy + 1, w, h), prev3 = ldu(A.objectid1, pvx + 1, pvy + 1, w, h);
        if (fabsf(motionInRange) < 0.1f) prev0 = prev1 = prev2 = prev3 = cur;
        float2 pos = make_float2((float)x + motion.x, (float)y + motion.y);
        float2 subpixel = make_float2(pos.x - floorf(pos.x), pos.y - floorf(pos.y));
        float3 historyDiffuse = FilterHistoryColor(A.prev_diffuse, w, h, subpixel, pvx, pvy, cur, prev0, prev1, prev2, prev3, spatialDiffuseSample);
        float3 historySpecular = FilterHistoryColor(A.prev_specular, w, h, subpixel, pvx, pvy, cur, prev0, prev1, prev2, prev3, spatialSpecularSample);
        float3 historyAlbedo = FilterHistoryColor(A.prev_albedo, w, h, subpixel, pvx, pvy, cur, prev0, prev1, prev2, prev3, currAlbedoSample);
        float currKeep = clampf(1.0f / (float)max(1, A.temporal_frames), 0.0f, 1.0f);
        float3 d = lerp3(historyDiffuse, xyz(currDiffuseSample), currKeep);
        float3 s = lerp3(historySpecular, xyz(currSpecularSample), currKeep);
        float3 a = lerp3(historyAlbedo, xyz(currAlbedoSample), currKeep);
        store_accum(A, idx, y, f4(d.x, d.y, d.z, load_h4(A.prev_diffuse, idx).w), f4(s.x, s.y, s.z, load_h4(A.prev_specular, idx).w),
                    f4(a.x, a.y, a.z, load_h4(A.prev_albedo, idx).w));
    } else {
        store_accum(A, idx, y, currDiffuseSample, currSpecularSample, currAlbedoSample);
    }
}

// ------------------------------------------------------------------------------------------------ DenoiseJBF
constexpr int DBX = 32, DBY = 8, DR = 5;
constexpr int DTW = DBX + 2 * DR, DTH = DBY + 2 * DR;
constexpr int DRP = 6, DTWP = DBX + 2 * DRP;  // TMA box: 6-column halo so that the box starts 16-byte aligned

struct DenoiseArgs {
    int w, h;
    const uint2 *accum_diffuse, *accum_specular, *accum_albedo;
    uint2* denoised;
    const GkUniformBufferObject* cam;
    int y_begin, y_end;  // rows processed (strip-parallel post: this rank's strip; `denoised` may then be rank 0's image over peer memory)
};

GK_DEV float exp2f_approx(float x) { float y; asm("ex2.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// Const_Func.slang:51-68, 82-126
GK_DEV float W_f(float x, float e0, float e1) {
    if (x <= e0) return 0;
    if (x >= e1) return 1;
    float a = (x - e0) / (e1 - e0);
    return a * a * (3 - 2 * a);
}
GK_DEV float H_f(float x, float e0, float e1) {
    if (x <= e0) return 0;
    if (x >= e1) return 1;
    return (x - e0) / (e1 - e0);
}
// FAST (GKCU_FAST_STENCIL=1, gkcu_set_fast_stencil): the hardware's approximate transcendentals (lg2.approx / ex2.approx, 2 ulp) and reciprocal
// instead of the pinned software exp / pow and the IEEE division. Slang leaves these intrinsics to the driver's compiler, so no reference bit
// pattern exists for them; the pinned build is bit-equal to the repo's oracle, the fast build is within the tolerance tests/ states.
template <bool FAST> GK_DEV float post_pow(float x, float y) { return FAST ? __powf(x, y) : gk_det_powf(x, y); }
template <bool FAST> GK_DEV float post_exp(float x) { return FAST ? __expf(x) : gk_det_expf(x); }
template <bool FAST> GK_DEV float post_div(float a, float b) { return FAST ? __fdividef(a, b) : a / b; }
template <bool FAST>
GK_DEV float GranTurismoTonemapper(float x) {
    const float e = 2.71828f;
    float P = 1, a = 0.7f, m = 0.22f, l = 0.4f, c = 1.33f, b = 0;
    float l0 = (P - m) * l / a;
    float L_x = m + a * (x - m);
    float T_x = m * post_pow<FAST>(x / m, c) + b;
    float S0 = m + l0;
    float S1 = m + a * l0;
    float C2 = a * P / (P - S1);
    float S_x = P - (P - S1) * post_pow<FAST>(e, -(C2 * (x - S0) / P));
    float w0_x = 1 - W_f(x, 0, m);
    float w2_x = H_f(x, m + l0, m + l0);
    float w1_x = 1 - w0_x - w2_x;
    return T_x * w0_x + L_x * w1_x + S_x * w2_x;
}
template <bool FAST>
GK_DEV float3 GT_Tonemapping(float3 x) { return f3(GranTurismoTonemapper<FAST>(x.x), GranTurismoTonemapper<FAST>(x.y), GranTurismoTonemapper<FAST>(x.z)); }
template <bool FAST>
GK_DEV float ST2084(float lin) {
    const float m1 = 0.1593017578125f, m2 = 78.84375f, c1 = 0.8359375f, c2 = 18.8515625f, c3 = 18.6875f, C = 10000.f;
    float L = lin / C;
    float Lm = post_pow<FAST>(L, m1);
    float N1 = (c1 + c2 * Lm), N2 = (1.0f + c3 * Lm);
    float N = N1 * (1.0f / N2);
    return post_pow<FAST>(N, m2);
}

// The spatial weight of a tap depends on its offset and BFSigma only: the 36 values are computed once per block (same expression, same
// bits as the shader's per-pixel evaluation, Process.DenoiseJBF.comp.slang:139) instead of 36 exp per pixel.
constexpr int DTAPS = (DR + 1) * (DR + 1);  // taps at odd offsets -DR..DR in both directions: 6 x 6
template <bool FAST>
GK_DEV void jbf_spatial_table(float* s_F, int tid, float sigma) {
    if (tid < DTAPS) {
        const int i = -DR + 2 * (tid / (DR + 1)), j = -DR + 2 * (tid % (DR + 1));
        const float dist = clampf(float(i * i + j * j) / float(DR * DR), 0.0f, 1.0f);
        s_F[tid] = gk_det_expf(-dist * dist / (2.0f * sigma * sigma));
    }
}
// :128-146 for one pixel of the staged tile
template <bool FAST>
GK_DEV float3 jbf_filter(const float4 (*s_diff)[DTW], const float* s_F, int ly, int lx, float sigmaL) {
    const float CenterLum = s_diff[ly][lx].w;
    float3 Total = f3(0, 0, 0);
    float Weight = 0;
    const float kL = FAST ? -1.4426950408889634f / (2.0f * sigmaL * sigmaL) : 0.0f;
#pragma unroll
    for (int i = -DR; i <= DR; i += 2) {
#pragma unroll
        for (int j = -DR; j <= DR; j += 2) {  // the centre pixel is not a tap (:133-141)
            const float4 t = s_diff[ly + i][lx + j];
            const float3 Ci = xyz(t);
            const float lumi = t.w;
            const float dlum = (CenterLum - lumi) * (CenterLum - lumi);
            const float Fi = s_F[((i + DR) / 2) * (DR + 1) + (j + DR) / 2];
            const float Li = FAST ? exp2f_approx(dlum * dlum * kL) : gk_det_expf(-dlum * dlum / (2.0f * sigmaL * sigmaL));
            Total = Total + Ci * Fi * Li;
            Weight += Fi * Li;
        }
    }
    if (FAST) {
        const float inv = __frcp_rn(Weight);
        return Total * inv;
    }
    return f3(Total.x / Weight, Total.y / Weight, Total.z / Weight);
}
// :150-180: tone mapping + store
template <bool FAST>
GK_DEV void compose_store(const GkUniformBufferObject* C, uint2* denoised, size_t idx, float3 Total) {
    if (C->HDR) {
        Total = f3(Total.x / 2000.0f, Total.y / 2000.0f, Total.z / 2000.0f);
        Total = GT_Tonemapping<FAST>(Total);
        Total = Total * 2000.0f;
        float3 pw = Total * C->PaperWhiteNit;
        store_h4(denoised, idx, f4(ST2084<FAST>(pw.x / 230.0f), ST2084<FAST>(pw.y / 230.0f), ST2084<FAST>(pw.z / 230.0f), 1.0f));
    } else {
        float3 pw = Total * C->PaperWhiteNit;
        float3 t = GT_Tonemapping<FAST>(f3(pw.x / 40000.0f, pw.y / 40000.0f, pw.z / 40000.0f));
        store_h4(denoised, idx, f4(t.x, t.y, t.z, 1.0f));
    }
}

template <bool FAST>
__global__ void __launch_bounds__(DBX* DBY) k_denoise(DenoiseArgs A) {
    __shared__ float4 s_diff[DTH][DTW];  // rgb + 0.001 bias, luminance in w
    __shared__ float s_F[DTAPS];
    const GkUniformBufferObject* C = A.cam;
    const int w = A.w, h = A.h;
    const int gx0 = (int)C->ViewportRect.x, gy0 = (int)C->ViewportRect.y;
    const int bx0 = blockIdx.x * DBX + gx0, by0 = blockIdx.y * DBY + gy0 + A.y_begin;
    const int tid = threadIdx.y * DBX + threadIdx.x;
    const int x = bx0 + threadIdx.x, y = by0 + threadIdx.y;
    const bool filter = C->BFSize > 0;
    if (filter) {
        jbf_spatial_table<FAST>(s_F, tid, C->BFSigma);
        for (int i = tid; i < DTW * DTH; i += DBX * DBY) {
            int ty = i / DTW, tx = i % DTW;
            float4 v = ld4(A.accum_diffuse, bx0 + tx - DR, by0 + ty - DR, w, h);
            float3 c = xyz(v) + f3(0.001f, 0.001f, 0.001f);  // FetchDiffuse, :14-17
            s_diff[ty][tx] = f4(c.x, c.y, c.z, dot(c, f3(0.212671f, 0.715160f, 0.072169f)));
        }
        __syncthreads();
    }
    if (!inb(x, y, w, h) || y >= A.y_end) return;
    const size_t idx = (size_t)y * w + x;
    float3 Total = f3(0, 0, 0);
    if (filter) {
        const float3 CenterSpec = xyz(load_h4(A.accum_specular, idx)) + f3(0.001f, 0.001f, 0.001f);
        Total = jbf_filter<FAST>(s_diff, s_F, threadIdx.y + DR, threadIdx.x + DR, C->BFSigmaLum * 100.0f);
        if (!C->DebugDraw_Lighting) Total = Total * xyz(load_h4(A.accum_albedo, idx)) + CenterSpec;
    } else {
        float3 d = xyz(load_h4(A.accum_diffuse, idx)), s = xyz(load_h4(A.accum_specular, idx));
        if (C->DebugDraw_Lighting) Total = d * f3(0.5f, 0.5f, 0.5f) + s;
        else Total = d * xyz(load_h4(A.accum_albedo, idx)) + s;
    }
    compose_store<FAST>(C, A.denoised, idx, Total);
}


// DenoiseJBF with the 42x18 ACCUM_DIFFUSE tile staged by one TMA copy (zero fill outside the image).
template <bool FAST>
__global__ void __launch_bounds__(DBX* DBY) k_denoise_tma(const __grid_constant__ CUtensorMap map_diffuse, DenoiseArgs A) {
    __shared__ alignas(128) uint2 raw[DTH][DTWP];
    __shared__ float4 s_diff[DTH][DTW];  // rgb + 0.001 bias, luminance in w
    __shared__ float s_F[DTAPS];
    __shared__ uint64_t bar;
    const GkUniformBufferObject* C = A.cam;
    const int w = A.w, h = A.h;
    const int gx0 = (int)C->ViewportRect.x, gy0 = (int)C->ViewportRect.y;
    const int bx0 = blockIdx.x * DBX + gx0, by0 = blockIdx.y * DBY + gy0 + A.y_begin;
    const int tid = threadIdx.y * DBX + threadIdx.x;
    const int x = bx0 + threadIdx.x, y = by0 + threadIdx.y;
    if (tid == 0) gktma::mbar_init(&bar, 1);
    __syncthreads();
    if (tid == 0) {
        gktma::mbar_expect_tx(&bar, (uint32_t)sizeof(raw));
        gktma::tma_load_2d(raw, &map_diffuse, bx0 - DRP, by0 - DR, &bar);
    }
    jbf_spatial_table<FAST>(s_F, tid, C->BFSigma);  // while the tile is in flight
    gktma::mbar_wait(&bar, 0);
    for (int i = tid; i < DTW * DTH; i += DBX * DBY) {
        int ty = i / DTW, tx = i % DTW;
        uint2 v = raw[ty][tx + (DRP - DR)];
        __half2 a = *reinterpret_cast<__half2*>(&v.x), b = *reinterpret_cast<__half2*>(&v.y);
        float2 fa = __half22float2(a), fb = __half22float2(b);
        float3 c = f3(fa.x, fa.y, fb.x) + f3(0.001f, 0.001f, 0.001f);  // FetchDiffuse, :14-17
        s_diff[ty][tx] = f4(c.x, c.y, c.z, dot(c, f3(0.212671f, 0.715160f, 0.072169f)));
    }
    __syncthreads();
    if (!inb(x, y, w, h) || y >= A.y_end) return;
    const size_t idx = (size_t)y * w + x;
    const float3 CenterSpec = xyz(load_h4(A.accum_specular, idx)) + f3(0.001f, 0.001f, 0.001f);
    float3 Total = jbf_filter<FAST>(s_diff, s_F, threadIdx.y + DR, threadIdx.x + DR, C->BFSigmaLum * 100.0f);
    if (!C->DebugDraw_Lighting) Total = Total * xyz(load_h4(A.accum_albedo, idx)) + CenterSpec;
    compose_store<FAST>(C, A.denoised, idx, Total);
}

}  // namespace

// CUDA loads a kernel's code at its first launch (lazy loading), and that load waits for the device to drain. In a multi-GPU frame the first
// launch of these kernels would come while this rank's barrier kernel is spinning for its peers - with several ranks in one process (the tests)
// that is a deadlock until the barrier times out. Resolve them up front.
void post_preload() {
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, k_reproject);
    cudaFuncGetAttributes(&fa, k_reproject_tma<0>);
    cudaFuncGetAttributes(&fa, k_reproject_tma<2>);
    cudaFuncGetAttributes(&fa, k_denoise<false>);
    cudaFuncGetAttributes(&fa, k_denoise<true>);
    cudaFuncGetAttributes(&fa, k_denoise_tma<false>);
    cudaFuncGetAttributes(&fa, k_denoise_tma<true>);
}

int run_reproject(Context& c, const GkcuFrameParams& p, const PostStrip* strip) {
    RenderTargets& t = c.rt;
    ReprojArgs A{};
    A.w = (int)t.w; A.h = (int)t.h;
    A.single_diffuse = t.single_diffuse; A.single_specular = t.single_specular; A.albedo = t.albedo; A.normal = t.normal;
    A.prev_diffuse = t.accum_diffuse[t.cur ^ 1]; A.prev_specular = t.accum_specular[t.cur ^ 1]; A.prev_albedo = t.accum_albedo[t.cur ^ 1];
    A.accum_diffuse = t.accum_diffuse[t.cur]; A.accum_specular = t.accum_specular[t.cur]; A.accum_albedo = t.accum_albedo[t.cur];
    A.objectid0 = t.objectid[t.cur]; A.objectid1 = t.objectid[t.cur ^ 1];
    A.motion = t.motion; A.moment = t.moment;
    A.progressive = p.progressive; A.fast_reproject = p.fast_reproject; A.temporal_frames = p.temporal_frames;
    A.y_begin = 0; A.y_end = (int)t.h; A.own_y0 = A.own_y1 = 0; A.n_mirror = 0;
    if (strip) {
        A.y_begin = std::max(0, strip->y_begin); A.y_end = std::min((int)t.h, strip->y_end);
        A.own_y0 = strip->own_y0; A.own_y1 = strip->own_y1; A.n_mirror = strip->n_mirror;
        for (int r = 0; r < strip->n_mirror; ++r) { A.mirror_diffuse[r] = strip->mirror_diffuse[r]; A.mirror_specular[r] = strip->mirror_specular[r]; A.mirror_albedo[r] = strip->mirror_albedo[r]; }
        if (A.y_end <= A.y_begin) return 0;
    }
    dim3 grid((t.w + RBX - 1) / RBX, (A.y_end - A.y_begin + RBY - 1) / RBY), block(RBX, RBY);
    // TMA staging needs 16-byte aligned rows (even width for the 8-byte pixels, width % 4 == 0 for the 4-byte ids)
    ReprojMaps M;
    const bool tma = (c.use_tma == 1 || c.use_tma == 2) && !p.progressive && (t.w % 4 == 0) &&
                     gktma::encode_2d(&M.diff, t.single_diffuse, CU_TENSOR_MAP_DATA_TYPE_UINT64, 8, t.w, t.h, RTW, RTH) &&
                     gktma::encode_2d(&M.spec, t.single_specular, CU_TENSOR_MAP_DATA_TYPE_UINT64, 8, t.w, t.h, RTW, RTH) &&
                     gktma::encode_2d(&M.norm, t.normal, CU_TENSOR_MAP_DATA_TYPE_UINT64, 8, t.w, t.h, RTW, RTH) &&
                     gktma::encode_2d(&M.ids, t.objectid[t.cur], CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, t.w, t.h, RIW, RTH) &&
                     gktma::encode_2d(&M.motion, t.motion, CU_TENSOR_MAP_DATA_TYPE_UINT64, 8, t.w, t.h, RMW, RMHH);
    if (tma) {
        if (p.fast_reproject) k_reproject_tma<0><<<grid, block, 0, c.stream>>>(M, A);
        else k_reproject_tma<2><<<grid, block, 0, c.stream>>>(M, A);
    } else {
        k_reproject<<<grid, block, 0, c.stream>>>(A);
    }
    c.launches++;
    GK_CUDA(cudaGetLastError());
    return 0;
}

int run_denoise(Context& c, const GkUniformBufferObject& ubo, const PostStrip* strip) {
    if (c.copy_pending) {  // an asynchronous present of the previous frame still reads RT_DENOISED
        cudaStreamWaitEvent(c.stream, c.copy_done, 0);
        c.copy_pending = false;
    }
    RenderTargets& t = c.rt;
    DenoiseArgs A;
    A.w = (int)t.w; A.h = (int)t.h;
    A.accum_diffuse = t.accum_diffuse[t.cur]; A.accum_specular = t.accum_specular[t.cur]; A.accum_albedo = t.accum_albedo[t.cur];
    A.denoised = t.denoised_out();
    A.cam = c.ubo.as<GkUniformBufferObject>();
    A.y_begin = 0; A.y_end = (int)t.h;
    if (strip) {
        A.y_begin = std::max(0, strip->own_y0); A.y_end = std::min((int)t.h, strip->own_y1);
        if (strip->denoised) A.denoised = strip->denoised;
        if (A.y_end <= A.y_begin) return 0;
    }
    dim3 grid((t.w + DBX - 1) / DBX, (A.y_end - A.y_begin + DBY - 1) / DBY), block(DBX, DBY);
    CUtensorMap map;
    if ((c.use_tma == 1 || c.use_tma == 3) && ubo.BFSize > 0 && (t.w % 2 == 0) && ((int)ubo.ViewportRect.x % 2 == 0) && gktma::encode_2d(&map, t.accum_diffuse[t.cur], CU_TENSOR_MAP_DATA_TYPE_UINT64, 8, t.w, t.h, DTWP, DTH))
        c.fast_stencil ? k_denoise_tma<true><<<grid, block, 0, c.stream>>>(map, A) : k_denoise_tma<false><<<grid, block, 0, c.stream>>>(map, A);
    else
        c.fast_stencil ? k_denoise<true><<<grid, block, 0, c.stream>>>(A) : k_denoise<false><<<grid, block, 0, c.stream>>>(A);
    c.launches++;
    GK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace gkcu
