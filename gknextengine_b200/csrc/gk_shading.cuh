// gk_shading.cuh — scene view and the shading helpers shared by the path tracer (gk_pathtrace.cu) and the probe bake
// (gk_probes.cu): material / texture / sky fetches, the ambient-cube probe lookup, hit-attribute reconstruction.
// Restates assets/shaders/common/{Shading.slang:146-199, 801-823; AmbientCube.slang:58-382; GeneralFunc.slang; BasicTypes.slang}.
#pragma once
#include "gk_bvh.cuh"
#include "gk_context.h"

namespace gkcu {
using namespace gkd;

struct TexDesc { const float4* texels; uint32_t w, h; uint32_t pad; };

struct SceneView {
    TraceScene ts;
    const uint2* vertices;  // GkGPUVertex as 3 x uint2
    const uint32_t* indices;
    const GkModelData* offsets;
    const GkMaterial* materials;
    uint32_t material_count;
    const GkNodeProxy* nodes;
    const uint4* shade_tris;  // 2 x uint4 per primitive
    const uint4* shade_inst;  // 2 x uint4 per instance
    const TexDesc* textures;
    uint32_t tex_count;
    const GkAmbientCube* cubes;
    const GkVoxelData* voxels;
    const GkUniformBufferObject* cam;
    const GkLightObject* lights;      // GPUScene.Lights (BasicTypes.slang:171-183)
    const GkSphericalHarmonics* sky_sh;  // GPUScene.HDRSHs, indexed like the textures (BasicTypes.slang:164-169)
    uint32_t sky_sh_count;
};


SceneView make_scene_view(const Context& c);  // gk_pathtrace.cu

namespace {

GK_DEV GkMaterial get_material(const SceneView& S, uint32_t idx) {  // BasicTypes.slang:376-379
    int i = min(GK_MAX_MATERIALS - 1, max(0, (int)idx));
    if (i >= (int)S.material_count) {
        GkMaterial z;
        z.Diffuse = {0, 0, 0, 0}; z.DiffuseTextureId = 0; z.MRATextureId = 0; z.NormalTextureId = 0; z.Fuzziness = 0; z.RefractionIndex = 0;
        z.MaterialModel = 0; z.Metalness = 0; z.RefractionIndex2 = 0; z.NormalTextureScale = 0; z.Reserverd2 = z.Reserverd3 = z.Reserverd4 = 0;
        return z;
    }
    const float4* p = reinterpret_cast<const float4*>(S.materials + i);
    float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
    GkMaterial m;
    m.Diffuse = {a.x, a.y, a.z, a.w};
    m.DiffuseTextureId = __float_as_int(b.x); m.MRATextureId = __float_as_int(b.y); m.NormalTextureId = __float_as_int(b.z); m.Fuzziness = b.w;
    m.RefractionIndex = c.x; m.MaterialModel = __float_as_uint(c.y); m.Metalness = c.z; m.RefractionIndex2 = c.w;
    m.NormalTextureScale = 0; m.Reserverd2 = m.Reserverd3 = m.Reserverd4 = 0;
    return m;
}

struct VtxLite { float3 pos, nrm; float2 uv; float4 tan; uint32_t matSlot; };
GK_DEV VtxLite unpack_vertex(const SceneView& S, uint32_t vi) {  // UnpackVertexV2, Const_Func.slang:356-368
    const uint2* p = S.vertices + (size_t)vi * 3;
    uint2 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
    VtxLite v;
    v.pos = f3(h2f(a.x & 0xFFFF), h2f(a.x >> 16), h2f(a.y & 0xFFFF));
    v.nrm = f3(h2f(b.x & 0xFFFF), h2f(b.x >> 16), h2f(b.y & 0xFFFF));
    v.uv = make_float2(h2f(a.y >> 16), h2f(b.y >> 16));
    uint32_t tw = c.y >> 16;
    v.tan = f4(h2f(c.x & 0xFFFF), h2f(c.x >> 16), h2f(c.y & 0xFFFF), (float)((tw >> 8) & 0xFF));
    v.matSlot = tw & 0xFF;
    return v;
}
GK_DEV uint32_t fetch_material_id(const GkNodeProxy* node, uint32_t slot) { return slot < 16 ? node->matId[slot] : 0; }  // :370-424

// Sampler2D.SampleLevel(uv, 0): linear, repeat (src/Vulkan/Sampler.hpp:9-26), full-precision weights
GK_DEV float4 sample_texture(const SceneView& S, uint32_t id, float2 uv) {
    if (id >= S.tex_count) return f4(0, 0, 0, 0);
    TexDesc tx = S.textures[id];
    if (tx.w == 0) return f4(0, 0, 0, 0);
    float x = uv.x * (float)tx.w - 0.5f, y = uv.y * (float)tx.h - 0.5f;
    float fx = floorf(x), fy = floorf(y);
    float ax = x - fx, ay = y - fy;
    int x0 = (int)fx, y0 = (int)fy;
    int xa = x0 % (int)tx.w; if (xa < 0) xa += tx.w;
    int xb = (x0 + 1) % (int)tx.w; if (xb < 0) xb += tx.w;
    int ya = y0 % (int)tx.h; if (ya < 0) ya += tx.h;
    int yb = (y0 + 1) % (int)tx.h; if (yb < 0) yb += tx.h;
    float4 t00 = __ldg(tx.texels + (size_t)ya * tx.w + xa), t10 = __ldg(tx.texels + (size_t)ya * tx.w + xb);
    float4 t01 = __ldg(tx.texels + (size_t)yb * tx.w + xa), t11 = __ldg(tx.texels + (size_t)yb * tx.w + xb);
    return lerp4(lerp4(t00, t10, ax), lerp4(t01, t11, ax), ay);
}

GK_DEV float4 SampleIBLV2(const SceneView& S, uint32_t skyIdx, float3 direction, float rotate) {  // Shading.slang:194-199
    float3 d = normalize(direction);
    float2 t = make_float2((atan2f(d.x, d.z) + GK_M_PI * rotate) * GK_M_1_OVER_TWO_PI, acosf(d.y) * GK_M_1_PI);
    float4 s = sample_texture(S, skyIdx, t);
    return f4(fminf(10.f, s.x), fminf(10.f, s.y), fminf(10.f, s.z), fminf(1.f, s.w));
}
GK_DEV float4 sky_color(const SceneView& S, float3 dir) {
    const GkUniformBufferObject* C = S.cam;
    return C->HasSky ? SampleIBLV2(S, C->SkyIdx, dir, C->SkyRotation) * C->SkyIntensity : f4(0, 0, 0, 0);
}

// AmbientCube.slang:71-78, 178-223, 290-382
// a / b for a constant b and its rounded reciprocal rb: one multiply and two FMAs (Markstein's correction) instead of the IEEE division
// sequence. Bit-identical to a / b for the operands used here - checked exhaustively for a = 0..1023, b = 1023 and a = 0..255, b = 255
// (tests/test_detmath.py) - which is what keeps the probe lookup equal to the oracle's plain divisions.
GK_DEV float div_const(float a, float b, float rb) {
    const float q = a * rb;
    return __fmaf_rn(__fmaf_rn(-q, b, a), rb, q);
}
GK_DEV float4 unpackRGB10A2(uint32_t p) {
    const float rb = 1.0f / 1023.0f;
    const float r = div_const(float(p & 0x3FF), 1023.0f, rb), g = div_const(float((p >> 10) & 0x3FF), 1023.0f, rb), b = div_const(float((p >> 20) & 0x3FF), 1023.0f, rb);
    return f4(r, g, b, 0.0f) * 512.f;
}
GK_DEV float4 sampleAmbientCubeHL2_Full(const GkAmbientCube& cube, float3 normal) {
    float4 color = f4(0, 0, 0, 0);
    float skyvis = 0.0f;
    float wx = fmaxf(normal.x, 0.0f), wnx = fmaxf(-normal.x, 0.0f), wy = fmaxf(normal.y, 0.0f), wny = fmaxf(-normal.y, 0.0f),
          wz = fmaxf(normal.z, 0.0f), wnz = fmaxf(-normal.z, 0.0f);
    float sum = wx + wnx + wy + wny + wz + wnz;
    uint32_t s0 = cube.skyVisibility_pznzpyny, s1 = cube.skyVisibility_pxnxs0s1;
    skyvis += float(s0 & 0xFF) * wz;
    skyvis += float((s0 >> 8) & 0xFF) * wnz;
    skyvis += float((s0 >> 16) & 0xFF) * wy;
    skyvis += float((s0 >> 24) & 0xFF) * wny;
    skyvis += float(s1 & 0xFF) * wx;
    skyvis += float((s1 >> 8) & 0xFF) * wnx;
    color = color + unpackRGB10A2(cube.PosX_D) * wx; color = color + unpackRGB10A2(cube.NegX_D) * wnx;
    color = color + unpackRGB10A2(cube.PosY_D) * wy; color = color + unpackRGB10A2(cube.NegY_D) * wny;
    color = color + unpackRGB10A2(cube.PosZ_D) * wz; color = color + unpackRGB10A2(cube.NegZ_D) * wnz;
    color = color + unpackRGB10A2(cube.PosX) * wx; color = color + unpackRGB10A2(cube.NegX) * wnx;
    color = color + unpackRGB10A2(cube.PosY) * wy; color = color + unpackRGB10A2(cube.NegY) * wny;
    color = color + unpackRGB10A2(cube.PosZ) * wz; color = color + unpackRGB10A2(cube.NegZ) * wnz;
    float k = (sum > 0.0f) ? (1.0f / sum) : 1.0f;
    color.x *= k; color.y *= k; color.z *= k;
    skyvis *= k;
    color.w = skyvis / 255.0f;
    return color;
}
GK_DEV float4 sampleAmbientCubeHL2_DI(const GkAmbientCube& cube, float3 normal) {  // AmbientCube.slang:128-151
    float4 color = f4(0, 0, 0, 0);
    float wx = fmaxf(normal.x, 0.0f), wnx = fmaxf(-normal.x, 0.0f), wy = fmaxf(normal.y, 0.0f), wny = fmaxf(-normal.y, 0.0f),
          wz = fmaxf(normal.z, 0.0f), wnz = fmaxf(-normal.z, 0.0f);
    float sum = wx + wnx + wy + wny + wz + wnz;
    color = color + unpackRGB10A2(cube.PosX_D) * wx; color = color + unpackRGB10A2(cube.NegX_D) * wnx;
    color = color + unpackRGB10A2(cube.PosY_D) * wy; color = color + unpackRGB10A2(cube.NegY_D) * wny;
    color = color + unpackRGB10A2(cube.PosZ_D) * wz; color = color + unpackRGB10A2(cube.NegZ_D) * wnz;
    return color * ((sum > 0.0f) ? (1.0f / sum) : 1.0f);
}
struct FullAmbientCubeSampler { GK_DEV static float4 Sample(const GkAmbientCube& c, float3 n) { return sampleAmbientCubeHL2_Full(c, n); } };
struct DIAmbientCubeSampler { GK_DEV static float4 Sample(const GkAmbientCube& c, float3 n) { return sampleAmbientCubeHL2_DI(c, n); } };

// Where a probe is read from. The bake updates a contiguous range of probes in place while its own rays sample probes
// (Bake.HwAmbientCube.comp.slang -> FaceTask -> interpolateAmbientCubesV2): in the reference that is a data race inside one
// dispatch. Here the range being baked is read from a snapshot taken before the launch, which picks one of the outcomes the
// race allows (every neighbour seen in its pre-dispatch state) and makes the result deterministic.
struct LiveCubes {
    const GkAmbientCube* cubes;
    GK_DEV GkAmbientCube fetch(int idx) const {  // read-only while the path tracer runs: seven 8-byte loads through the read-only path
        const uint2* p = reinterpret_cast<const uint2*>(cubes + idx);
        const uint2 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2), d = __ldg(p + 3), e = __ldg(p + 4), f = __ldg(p + 5), g = __ldg(p + 6);
        GkAmbientCube r;
        r.PosZ = a.x; r.NegZ = a.y; r.PosY = b.x; r.NegY = b.y; r.PosX = c.x; r.NegX = c.y;
        r.PosZ_D = d.x; r.NegZ_D = d.y; r.PosY_D = e.x; r.NegY_D = e.y; r.PosX_D = f.x; r.NegX_D = f.y;
        r.skyVisibility_pznzpyny = g.x; r.skyVisibility_pxnxs0s1 = g.y;
        return r;
    }
};
struct SnapshotCubes {
    const GkAmbientCube* cubes;
    const GkAmbientCube* snapshot;  // copies of cubes[first .. first + count)
    uint32_t first, count;
    GK_DEV GkAmbientCube fetch(int idx) const {
        const uint32_t k = (uint32_t)idx - first;
        return k < count ? snapshot[k] : cubes[idx];
    }
};

template <class Sampler, class Fetch>
__device__ __noinline__ float4 interpolateAmbientCubesV2(const SceneView& S, const Fetch& F, float3 inPos, float3 normal) {
    const float3 CUBE_OFFSET = f3(-GK_CUBE_SIZE_XY / 2, -1.375f, -GK_CUBE_SIZE_XY / 2) * 0.25f;
    float3 nearpos = (inPos - CUBE_OFFSET) * (1.0f / 0.25f);
    if (nearpos.x < 0 || nearpos.y < 0 || nearpos.z < 0 || nearpos.x > GK_CUBE_SIZE_XY - 2 || nearpos.y > GK_CUBE_SIZE_Z - 2 ||
        nearpos.z > GK_CUBE_SIZE_XY - 2)
        return f4(0, 0, 0, 1);
    if (!S.cubes) return f4(0, 0, 0, 0);  // zero probe state: every probe is skipped at AmbientCube.slang:327
    int bx = (int)floorf(nearpos.x), by = (int)floorf(nearpos.y), bz = (int)floorf(nearpos.z);
    float3 fr = f3(nearpos.x - floorf(nearpos.x), nearpos.y - floorf(nearpos.y), nearpos.z - floorf(nearpos.z));
    float totalWeight = 0.0f;
    float4 result = f4(0, 0, 0, 0);
    // the eight voxels' distance words first (independent 8-byte loads in flight together), then the probes that pass
    uint2 vox[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int idx = (by + ((i >> 1) & 1)) * GK_CUBE_SIZE_XY * GK_CUBE_SIZE_XY + (bz + ((i >> 2) & 1)) * GK_CUBE_SIZE_XY + (bx + (i & 1));
        vox[i] = reinterpret_cast<const uint2*>(S.voxels + idx)[1];
    }
    const float r255 = 1.0f / 255.0f;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        int ox = i & 1, oy = (i >> 1) & 1, oz = (i >> 2) & 1;
        int idx = (by + oy) * GK_CUBE_SIZE_XY * GK_CUBE_SIZE_XY + (bz + oz) * GK_CUBE_SIZE_XY + (bx + ox);
        const uint32_t u0 = vox[i].x, u1 = vox[i].y;
        float d0y = div_const(float((u0 >> 8) & 0xFF), 255.0f, r255), distPZ = div_const(float((u0 >> 16) & 0xFF), 255.0f, r255),
              distNZ = div_const(float((u0 >> 24) & 0xFF), 255.0f, r255);
        float distPX = div_const(float(u1 & 0xFF), 255.0f, r255), distNX = div_const(float((u1 >> 8) & 0xFF), 255.0f, r255),
              distPY = div_const(float((u1 >> 16) & 0xFF), 255.0f, r255), distNY = div_const(float((u1 >> 24) & 0xFF), 255.0f, r255);
        if (d0y < 0.01f) continue;
        float3 probeToLocal = fr - f3((float)ox, (float)oy, (float)oz);
        float distanceToProbe = length(probeToLocal);
        float3 dir = normalize(probeToLocal);
        float hitLength = sqrtf(fmaxf(dir.x, 0.0f)) * distPX + sqrtf(fmaxf(-dir.x, 0.0f)) * distNX + sqrtf(fmaxf(dir.y, 0.0f)) * distPY +
                          sqrtf(fmaxf(-dir.y, 0.0f)) * distNY + sqrtf(fmaxf(dir.z, 0.0f)) * distPZ + sqrtf(fmaxf(-dir.z, 0.0f)) * distNZ;
        if (distanceToProbe > hitLength + 0.05f) continue;
        float wx = ox == 0 ? (1.0f - fr.x) : fr.x, wy = oy == 0 ? (1.0f - fr.y) : fr.y, wz = oz == 0 ? (1.0f - fr.z) : fr.z;
        float weight = wx * wy * wz;
        result = result + Sampler::Sample(F.fetch(idx), normal) * weight;
        totalWeight += weight;
    }
    return totalWeight > 0.0f ? result / totalWeight : f4(0, 0, 0, 0);
}


// AmbientCube.slang:58-69, 112-119
GK_DEV uint32_t packRGB10A2(float4 color) {
    const float4 c = f4(clampf(color.x / 512.f, 0.0f, 1.0f), clampf(color.y / 512.f, 0.0f, 1.0f), clampf(color.z / 512.f, 0.0f, 1.0f),
                        clampf(color.w / 512.f, 0.0f, 1.0f));
    const uint32_t r = (uint32_t)(c.x * 1023.0f), g = (uint32_t)(c.y * 1023.0f), b = (uint32_t)(c.z * 1023.0f), a = (uint32_t)(c.w * 3.0f);
    return r | (g << 10) | (b << 20) | (a << 30);
}
GK_DEV uint32_t LerpPackedColorAlt(uint32_t c0, float4 c1, float t) { return packRGB10A2(lerp4(unpackRGB10A2(c0), c1, t)); }

// Shading.slang:146-192 (EvaluateSH + SampleIBLRough): 9-coefficient SH of the sky, rotated about Y.
GK_DEV float4 SampleIBLRough(const SceneView& S, uint32_t skyIdx, float3 normal, float rotate) {
    if (skyIdx >= S.sky_sh_count) return f4(0, 0, 0, 1);
    const float angle = (1.0f - rotate) * 3.14159265358979323846f;
    float sinAngle, cosAngle;
    gk_det_sincosf(angle, &sinAngle, &cosAngle);
    const float3 n = f3(normal.x * cosAngle + normal.z * sinAngle, normal.y, -normal.x * sinAngle + normal.z * cosAngle);
    const float SH_C0 = 0.282095f, SH_C1 = 0.488603f, SH_C2 = 1.092548f, SH_C3 = 0.315392f, SH_C4 = 0.546274f;
    float basis[9];
    basis[0] = SH_C0;
    basis[1] = -SH_C1 * n.y;
    basis[2] = SH_C1 * n.z;
    basis[3] = -SH_C1 * n.x;
    basis[4] = SH_C2 * n.x * n.y;
    basis[5] = -SH_C2 * n.y * n.z;
    basis[6] = SH_C3 * (3.f * n.y * n.y - 1.0f);
    basis[7] = -SH_C2 * n.x * n.z;
    basis[8] = SH_C4 * (n.x * n.x - n.z * n.z);
    const float* co = S.sky_sh[skyIdx].coefficients[0];
    float3 color = f3(0, 0, 0);
    for (int i = 0; i < 9; ++i) {
        color.x += co[i] * basis[i];
        color.y += co[9 + i] * basis[i];
        color.z += co[18 + i] * basis[i];
    }
    return f4(color.x, color.y, color.z, 1.0f);
}

// What FHardwareRayTracerV2::TraceRay returns for a committed hit (Shading.slang:801-823), from the compact shading records.
struct HitVertex { float3 pos, nrm; float2 uv; uint32_t mat; };
GK_DEV HitVertex resolve_hit(const SceneView& S, float3 ro, float3 rd, const Hit& h) {
    const uint4 si0 = __ldg(S.shade_inst + (size_t)h.inst * 2);
    const float4* wm = reinterpret_cast<const float4*>(S.ts.inst + h.inst);
    const float4 m0 = __ldg(wm), m1 = __ldg(wm + 1), m2 = __ldg(wm + 2);
    const uint4* tp = S.shade_tris + ((size_t)si0.x + h.prim) * 2;
    const uint4 t0 = __ldg(tp), t1 = __ldg(tp + 1);
    const float3 n0 = f3(h2f(t0.x & 0xFFFF), h2f(t0.x >> 16), h2f(t0.y & 0xFFFF));
    const float3 n1 = f3(h2f(t0.y >> 16), h2f(t0.z & 0xFFFF), h2f(t0.z >> 16));
    const float3 n2 = f3(h2f(t0.w & 0xFFFF), h2f(t0.w >> 16), h2f(t1.x & 0xFFFF));
    const float2 uv0 = make_float2(h2f(t1.x >> 16), h2f(t1.y & 0xFFFF)), uv1 = make_float2(h2f(t1.y >> 16), h2f(t1.z & 0xFFFF)),
                 uv2 = make_float2(h2f(t1.z >> 16), h2f(t1.w & 0xFFFF));
    const uint32_t slot = (t1.w >> 16) & 0xFF;
    HitVertex v;
    if (slot < 3) v.mat = slot == 0 ? si0.y : (slot == 1 ? si0.z : si0.w);
    else if (slot < 7) {
        const uint4 si1 = __ldg(S.shade_inst + (size_t)h.inst * 2 + 1);
        v.mat = slot == 3 ? si1.x : (slot == 4 ? si1.y : (slot == 5 ? si1.z : si1.w));
    } else v.mat = fetch_material_id(S.nodes + h.inst, slot);
    const float3 n = n0 + (n1 - n0) * h.u + (n2 - n0) * h.v;  // Mix(), PreProcessor.slang:24
    // mul(WorldToObject4x3, n) = transpose(w2o 3x3) * n
    const float3 nw = f3(m0.x * n.x + m1.x * n.y + m2.x * n.z, m0.y * n.x + m1.y * n.y + m2.y * n.z, m0.z * n.x + m1.z * n.y + m2.z * n.z);
    v.nrm = normalize(nw);
    v.uv = uv0 + (uv1 - uv0) * h.u + (uv2 - uv0) * h.v;
    v.pos = ro + rd * h.t;
    return v;
}

}  // namespace
}  // namespace gkcu
