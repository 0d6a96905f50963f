// gk_shading.cuh — scene view and the shading helpers shared by the path tracer (gk_pathtrace.cu) and the probe bake
// (gk_probes.cu): material / texture / sky fetches, the ambient-cube probe lookup, hit-attribute reconstruction.
// Restates assets/shaders/common/{Shading.slang:146-199, 801-823; AmbientCube.slang:58-382; GeneralFunc.slang; BasicTypes.slang}.
#pragma once
#include "gk_bvh.cuh"
#include "gk_context.h"

namespace gkcu {
using namespace gkd;

// Device layout of the probe grid. The reference keeps two arrays (AmbientCube 56 B, VoxelData 16 B, BasicTypes.slang:116-151); a lookup
// needs a probe's two distance words and, if it passes, its 14 colour / visibility words. Here they share one 64-byte aligned record
// (two 32-byte sectors of one line) and the words only the voxeliser and the bake touch (material, age) sit in a side array.
// gkcu_set_probes / gkcu_probes_read convert from and to the reference's two arrays.
struct __align__(64) ProbeRec {
    uint32_t cube[14];   // GkAmbientCube word for word
    uint32_t dist0;      // VoxelData.distanceToSolid_gg_z01
    uint32_t dist1;      // VoxelData.distanceToSolid_x01_y01
};
struct ProbeMeta { uint32_t matId, age; };
static_assert(sizeof(ProbeRec) == 64 && sizeof(GkAmbientCube) == 56, "probe record");

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
    const ProbeRec* probes;   // packed probe grid (nullptr: zero probe state)
    const ProbeMeta* probe_meta;
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

// v mod n in [0, n) (repeat addressing). Coordinates within one tile of the image - every sky lookup, most uv sets - take two
// compares instead of the integer division sequence (4 % of k_shade's instructions went into the four `%` of a bilinear fetch).
GK_DEV int wrap_texel(int v, int n) {
    if ((unsigned)(v + n) < 3u * (unsigned)n) return v < 0 ? v + n : (v >= n ? v - n : v);
    const int r = v % n;
    return r < 0 ? r + n : r;
}
// Sampler2D.SampleLevel(uv, 0): linear, repeat (src/Vulkan/Sampler.hpp:9-26), full-precision weights
GK_DEV float4 sample_texture(const SceneView& S, uint32_t id, float2 uv) {
    if (id >= S.tex_count) return f4(0, 0, 0, 0);
    TexDesc tx = S.textures[id];
    if (tx.w == 0) return f4(0, 0, 0, 0);
    float x = uv.x * (float)tx.w - 0.5f, y = uv.y * (float)tx.h - 0.5f;
    float fx = floorf(x), fy = floorf(y);
    float ax = x - fx, ay = y - fy;
    int x0 = (int)fx, y0 = (int)fy;
    const int xa = wrap_texel(x0, (int)tx.w), xb = wrap_texel(x0 + 1, (int)tx.w);
    const int ya = wrap_texel(y0, (int)tx.h), yb = wrap_texel(y0 + 1, (int)tx.h);
    float4 t00 = __ldg(tx.texels + (size_t)ya * tx.w + xa), t10 = __ldg(tx.texels + (size_t)ya * tx.w + xb);
    float4 t01 = __ldg(tx.texels + (size_t)yb * tx.w + xa), t11 = __ldg(tx.texels + (size_t)yb * tx.w + xb);
    return lerp4(lerp4(t00, t10, ax), lerp4(t01, t11, ax), ay);
}

GK_DEV float4 SampleIBLV2(const SceneView& S, uint32_t skyIdx, float3 direction, float rotate) {  // Shading.slang:194-199
    float3 d = normalize(direction);
    float2 t = make_float2((gk_det_atan2f(d.x, d.z) + GK_M_PI * rotate) * GK_M_1_OVER_TWO_PI, gk_det_acosf(d.y) * GK_M_1_PI);
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
GK_DEV GkAmbientCube cube_of(const uint4 a, const uint4 b, const uint4 c, const uint4 d) {
    GkAmbientCube r;
    r.PosZ = a.x; r.NegZ = a.y; r.PosY = a.z; r.NegY = a.w; r.PosX = b.x; r.NegX = b.y;
    r.PosZ_D = b.z; r.NegZ_D = b.w; r.PosY_D = c.x; r.NegY_D = c.y; r.PosX_D = c.z; r.NegX_D = c.w;
    r.skyVisibility_pznzpyny = d.x; r.skyVisibility_pxnxs0s1 = d.y;
    return r;
}
struct LiveCubes {
    const ProbeRec* probes;
    GK_DEV GkAmbientCube fetch(int idx) const {  // read-only while the path tracer runs: four 16-byte loads of one line, read-only path
        const uint4* p = reinterpret_cast<const uint4*>(probes + idx);
        return cube_of(__ldg(p), __ldg(p + 1), __ldg(p + 2), __ldg(p + 3));
    }
};
struct SnapshotCubes {
    const ProbeRec* probes;
    const ProbeRec* snapshot;  // copies of probes[first .. first + count)
    uint32_t first, count;
    GK_DEV GkAmbientCube fetch(int idx) const {
        const uint32_t k = (uint32_t)idx - first;
        const uint4* p = reinterpret_cast<const uint4*>(k < count ? snapshot + k : probes + idx);
        return cube_of(p[0], p[1], p[2], p[3]);
    }
};

// One of the eight probes around a point (AmbientCube.slang:318-376): its trilinear weight and weighted sample, or pass = false when the
// probe is inside geometry or hidden behind a surface.
template <class Sampler, class Fetch>
GK_DEV bool probe_term(const SceneView& S, const Fetch& F, int bx, int by, int bz, float3 fr, float3 normal, int i, float4& term, float& weight) {
    const int ox = i & 1, oy = (i >> 1) & 1, oz = (i >> 2) & 1;
    const int idx = (by + oy) * GK_CUBE_SIZE_XY * GK_CUBE_SIZE_XY + (bz + oz) * GK_CUBE_SIZE_XY + (bx + ox);
    const uint2 vox = reinterpret_cast<const uint2*>(S.probes + idx)[7];  // dist0, dist1: never written during a lookup or a bake
    const uint32_t u0 = vox.x, u1 = vox.y;
    const float r255 = 1.0f / 255.0f;
    const float d0y = div_const(float((u0 >> 8) & 0xFF), 255.0f, r255), distPZ = div_const(float((u0 >> 16) & 0xFF), 255.0f, r255),
                distNZ = div_const(float((u0 >> 24) & 0xFF), 255.0f, r255);
    const float distPX = div_const(float(u1 & 0xFF), 255.0f, r255), distNX = div_const(float((u1 >> 8) & 0xFF), 255.0f, r255),
                distPY = div_const(float((u1 >> 16) & 0xFF), 255.0f, r255), distNY = div_const(float((u1 >> 24) & 0xFF), 255.0f, r255);
    if (d0y < 0.01f) return false;
    const float3 probeToLocal = fr - f3((float)ox, (float)oy, (float)oz);
    const float distanceToProbe = length(probeToLocal);
    const float3 dir = normalize(probeToLocal);
    const float hitLength = sqrtf(fmaxf(dir.x, 0.0f)) * distPX + sqrtf(fmaxf(-dir.x, 0.0f)) * distNX + sqrtf(fmaxf(dir.y, 0.0f)) * distPY +
                            sqrtf(fmaxf(-dir.y, 0.0f)) * distNY + sqrtf(fmaxf(dir.z, 0.0f)) * distPZ + sqrtf(fmaxf(-dir.z, 0.0f)) * distNZ;
    if (distanceToProbe > hitLength + 0.05f) return false;
    const float wx = ox == 0 ? (1.0f - fr.x) : fr.x, wy = oy == 0 ? (1.0f - fr.y) : fr.y, wz = oz == 0 ? (1.0f - fr.z) : fr.z;
    weight = wx * wy * wz;
    term = Sampler::Sample(F.fetch(idx), normal) * weight;
    return true;
}

// Warp-cooperative interpolateAmbientCubesV2 (AmbientCube.slang:290-382). In the shade kernel only a couple of lanes per warp end their
// path in a given wave, so a per-lane lookup ran its 8 probes x ~350 instructions at 2 of 32 lanes (ncu, profiles/r01_experiments.md).
// Here the lanes that need a lookup are served four at a time: lane l evaluates probe (l & 7) of the (l >> 3)-th requester, and lane 0 of
// each group of eight adds the eight terms in probe order - the same order, and therefore the same rounding, as the reference's loop.
// Every lane of a converged warp must call this (need = false for the ones that only help); a warp that arrives incomplete falls back
// to per-lane evaluation.
template <class Sampler, class Fetch>
__device__ __noinline__ float4 interpolateAmbientCubesCoop(const SceneView& S, const Fetch& F, bool need, float3 inPos, float3 normal) {
    const float3 CUBE_OFFSET = f3(-GK_CUBE_SIZE_XY / 2, -1.375f, -GK_CUBE_SIZE_XY / 2) * 0.25f;
    const float3 nearpos = (inPos - CUBE_OFFSET) * (1.0f / 0.25f);
    float4 out = f4(0, 0, 0, 0);
    if (need && (nearpos.x < 0 || nearpos.y < 0 || nearpos.z < 0 || nearpos.x > GK_CUBE_SIZE_XY - 2 || nearpos.y > GK_CUBE_SIZE_Z - 2 ||
                 nearpos.z > GK_CUBE_SIZE_XY - 2)) {
        out = f4(0, 0, 0, 1);  // outside the probe grid
        need = false;
    }
    if (!S.probes) return out;  // zero probe state: every probe is skipped at :327
    const unsigned act = __activemask();
    if (act != 0xffffffffu) {  // incomplete warp (last block, or a caller inside divergent code): per-lane evaluation
        if (need) {
            const int bx = (int)floorf(nearpos.x), by = (int)floorf(nearpos.y), bz = (int)floorf(nearpos.z);
            const float3 fr = f3(nearpos.x - floorf(nearpos.x), nearpos.y - floorf(nearpos.y), nearpos.z - floorf(nearpos.z));
            float totalWeight = 0.0f;
            float4 result = f4(0, 0, 0, 0);
            for (int i = 0; i < 8; i++) {
                float4 term;
                float weight;
                if (!probe_term<Sampler>(S, F, bx, by, bz, fr, normal, i, term, weight)) continue;
                result = result + term;
                totalWeight += weight;
            }
            out = totalWeight > 0.0f ? result / totalWeight : f4(0, 0, 0, 0);
        }
        return out;
    }
    unsigned pending = __ballot_sync(0xffffffffu, need);
    const int lane = threadIdx.x & 31, g = lane >> 3, i = lane & 7;
    while (pending) {
        // the g-th requester still pending is this group's customer (none: the group idles this round)
        unsigned rem = pending;
        for (int k = 0; k < g; ++k) rem &= rem - 1;
        const int src = rem ? __ffs(rem) - 1 : -1;
        const int from = src < 0 ? lane : src;
        const float px = __shfl_sync(0xffffffffu, nearpos.x, from), py = __shfl_sync(0xffffffffu, nearpos.y, from), pz = __shfl_sync(0xffffffffu, nearpos.z, from);
        const float nx = __shfl_sync(0xffffffffu, normal.x, from), ny = __shfl_sync(0xffffffffu, normal.y, from), nz = __shfl_sync(0xffffffffu, normal.z, from);
        float4 term = f4(0, 0, 0, 0);
        float weight = 0.0f;
        bool pass = false;
        if (src >= 0) {
            const int bx = (int)floorf(px), by = (int)floorf(py), bz = (int)floorf(pz);
            const float3 fr = f3(px - floorf(px), py - floorf(py), pz - floorf(pz));
            pass = probe_term<Sampler>(S, F, bx, by, bz, fr, f3(nx, ny, nz), i, term, weight);
        }
        // lane 0 of the group adds the eight terms in probe order
        float4 result = f4(0, 0, 0, 0);
        float totalWeight = 0.0f;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int l = (lane & 24) + k;
            const float tx = __shfl_sync(0xffffffffu, term.x, l), ty = __shfl_sync(0xffffffffu, term.y, l), tz = __shfl_sync(0xffffffffu, term.z, l),
                        tw = __shfl_sync(0xffffffffu, term.w, l), w = __shfl_sync(0xffffffffu, weight, l);
            const bool p = __shfl_sync(0xffffffffu, pass ? 1 : 0, l) != 0;
            if (p) {
                result = result + f4(tx, ty, tz, tw);
                totalWeight += w;
            }
        }
        const float4 res = totalWeight > 0.0f ? result / totalWeight : f4(0, 0, 0, 0);
        // hand the four results to their requesters and retire them
        unsigned served = 0;
#pragma unroll
        for (int gg = 0; gg < 4; ++gg) {
            const int s_lane = __shfl_sync(0xffffffffu, src, gg * 8);
            const float rx = __shfl_sync(0xffffffffu, res.x, gg * 8), ry = __shfl_sync(0xffffffffu, res.y, gg * 8), rz = __shfl_sync(0xffffffffu, res.z, gg * 8),
                        rw = __shfl_sync(0xffffffffu, res.w, gg * 8);
            if (s_lane >= 0) {
                served |= 1u << s_lane;
                if (lane == s_lane) out = f4(rx, ry, rz, rw);
            }
        }
        pending &= ~served;
    }
    return out;
}

// Per-lane form (callers inside divergent code, e.g. the probe bake's hit branch).
template <class Sampler, class Fetch>
__device__ __noinline__ float4 interpolateAmbientCubesV2(const SceneView& S, const Fetch& F, float3 inPos, float3 normal) {
    const float3 CUBE_OFFSET = f3(-GK_CUBE_SIZE_XY / 2, -1.375f, -GK_CUBE_SIZE_XY / 2) * 0.25f;
    const float3 nearpos = (inPos - CUBE_OFFSET) * (1.0f / 0.25f);
    if (nearpos.x < 0 || nearpos.y < 0 || nearpos.z < 0 || nearpos.x > GK_CUBE_SIZE_XY - 2 || nearpos.y > GK_CUBE_SIZE_Z - 2 ||
        nearpos.z > GK_CUBE_SIZE_XY - 2)
        return f4(0, 0, 0, 1);
    if (!S.probes) return f4(0, 0, 0, 0);
    const int bx = (int)floorf(nearpos.x), by = (int)floorf(nearpos.y), bz = (int)floorf(nearpos.z);
    const float3 fr = f3(nearpos.x - floorf(nearpos.x), nearpos.y - floorf(nearpos.y), nearpos.z - floorf(nearpos.z));
    float totalWeight = 0.0f;
    float4 result = f4(0, 0, 0, 0);
    for (int i = 0; i < 8; i++) {
        float4 term;
        float weight;
        if (!probe_term<Sampler>(S, F, bx, by, bz, fr, normal, i, term, weight)) continue;
        result = result + term;
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
