// gk_pathtrace.cu — the "rt pass": primary visibility + G-buffer + the path tracer.
//
// Restates the reference's megakernel (assets/shaders/Core.PathTracing.comp.slang:5-31 ->
// FPathTracingRendererV2, assets/shaders/common/Shading.slang:995-1261) as a per-pixel state machine so the same
// shading code runs in two drivers:
//   * wavefront (default): k_primary -> { k_trace (persistent threads, dynamic ray fetch) -> k_shade (advance the
//     state machine, warp-ballot compaction of the survivors) }* -> k_tail (finishes the thin tail in one launch)
//   * megakernel (tracer = 1): k_primary -> k_tail over every pixel (one thread per pixel, traversal inlined).
// The reference threads ONE pcg4d state per pixel through all samples and bounces (Shading.slang:1232, 1117, 1130,
// 1135-1147, 1179); a pixel therefore has one path in flight at a time and samples run in order, which keeps the
// random stream - and with it every ray - identical to the megakernel order.
#include <algorithm>
#include <cstdio>
#include <type_traits>

#include <cub/device/device_radix_sort.cuh>

#ifndef GK_NODE_REPEAT
#define GK_NODE_REPEAT 1
#endif
#ifndef GK_TRACE_MINB
#define GK_TRACE_MINB 7  // resident blocks per SM k_trace<5> is compiled for and its carve-out is sized for (70 registers fit 7; compiled for 6 it ran 0.6 % slower on a carve-out one block short, 8 -> 64 registers is 2.4 % slower)
#endif
#include "gk_context.h"
#include "gk_shading.cuh"

using namespace gkd;

namespace gkcu {

struct RTView {
    uint32_t w, h;
    uint2 *single_diffuse, *single_specular, *albedo, *normal;
    uint32_t *objectid0, *primitiveid;
    float2* motion;
    float* prev_depth;
    uint16_t* moment;
    const uint32_t* vis;  // RT_MINIGBUFFER when the frame takes its primary visibility from the rasterised buffer, else nullptr
};

// Per-pixel path state (SoA, 16-byte records) + per-slot ray/hit records of the wavefront queues.
// "Hot" state (read and written on every path segment) lives per queue SLOT, double-buffered with the ray records, so k_shade
// reads and writes it coalesced without a pixel indirection. "Cold" state (per-pixel accumulators and the primary-hit record)
// is touched only at sample boundaries.
struct PathView {
    uint4* rng;                                          // hot, [2][pixels]
    float4 *pos_mat, *nrm, *dir, *ray_color;   // queue slots, [2][pixels]; pos_mat / nrm / dir only for any-hit slots
    float4 *final_color, *final_refl;                    // cold, per pixel: running sums of the finished samples
    float4 *prim_pos_mat, *prim_nrm_cnt, *prim_albedo;   // cold, per pixel: primary hit (position|material, normal|sample range, albedo)
    size_t pixels;
    float4 *ray_o, *ray_d;      // per slot: origin|tmin, dir|tmax
    uint32_t* slot_pixel[2];    // per slot: pixel | any-hit flag in bit 31 (ping-pong)
    float4* hit_tuv;            // per slot: t, u, v, inst
    uint32_t* hit_prim;
    uint32_t* counts;           // [0],[1] queue sizes (ping-pong), [2] trace cursor, [4..] ray counters
    float4 *radiance, *direct;  // multi-GPU: un-normalised sums (2 planes) and rank 0's direct term
    uint32_t* publish;          // mapped host word this wave's k_trace stores its input queue size to (nullptr: none)
    const uint32_t* order;      // k_trace: queue position -> slot (rays binned by origin cell and direction), nullptr = queue order
};

struct PathLanes { PathView v[4]; uint32_t n; };

struct PathBuffers {
    DevBuf rng, pos_mat, nrm_flags, dir_cnt, ray_color, final_color, final_refl, prim_pos_mat, prim_nrm_cnt, prim_albedo;
    DevBuf ray_o, ray_d, slot_pixel[2], hit_tuv, hit_prim, counts, radiance, direct;
    DevBuf sort_keys[2], sort_vals[2], sort_tmp;  // ray reordering (GKCU_SORT_RAYS)
    size_t sort_tmp_bytes = 0;
    DevBuf tail_order, tail_hist;  // GKCU_TAIL_SORT: tail slots ordered by the samples their pixel still has to run, and the 2 x 128 bucket counters per lane
    DevBuf pool_stack;  // k_trace_pool / k_trace_wpool: traversal stacks, [lane][block (, warp)][depth][slot]
    size_t pool_stack_stride = 0;  // uint2 entries per lane
    uint32_t* host_counts = nullptr;  // Context::pub_host + 64: mapped pinned memory, 64 words per lane, written by kernels
    size_t pixels = 0;
    size_t lane_cap = 0;  // queue capacity of one lane
    int lanes = 1;
};

struct FrameCfg {
    uint32_t frame_seed;     // TotalFrames, or TotalFrames * world + rank
    uint32_t rank, world;
    uint32_t to_radiance;    // 1: leave un-normalised sums in the radiance buffer (multi-GPU)
    uint32_t exit_after_first;
};

namespace {

#define ST_FIRST 0u
#define ST_BOUNCE 1u
#define ST_SUN 2u
#define ST_DIRECT 3u
#define FL_HIT_REFLECT 1u
#define FL_HIT_METAL 2u
#define FL_CHANCE_REFLECT 4u
#define FL_BACKFACE_ND 8u
#define FL_STAGE_SHIFT 4
#define FL_STAGE_MASK (3u << FL_STAGE_SHIFT)
#define FL_PRIM_DIELECTRIC 64u

#define REQ_DONE 0
#define REQ_CLOSEST 1
#define REQ_ANY 2

struct PathRegs {
    uint4 rng;
    float3 pos, nrm;   // vertexStart.Position / Normal
    uint32_t mat;      // vertexStart.MaterialIndex
    float3 dir;        // direction
    float4 rayColor;
    uint32_t flags;
    uint32_t sample, bounce, maxBounces;
    float3 ro, rd;     // pending ray
};

// ------------------------------------------------------------------------------------------------ primary hit
struct PrimVertex { float3 pos, nrm; float2 uv; uint32_t matSlot; float3 lnrm; float4 ltan; };

// Common.get_material_dataV2, GeneralFunc.slang:6-56 (the world-space tangent is not used on this path; the local normal and tangent feed the normal map)
GK_DEV PrimVertex get_material_dataV2(const SceneView& S, uint32_t nodeIdx, uint32_t prim, float3 ray_origin, float3 ray_direction) {
    const GkNodeProxy* proxy = S.nodes + nodeIdx;
    const GkModelData* model = S.offsets + proxy->modelId;
    const uint32_t vo = model->vertexOffset, io = model->indexOffset;
    float3 positions[3], normals[3], lnrm[3];
    float4 ltan[3];
    float2 tex[3];
    uint32_t matid = 0;
    for (int i = 0; i != 3; ++i) {
        VtxLite v = unpack_vertex(S, vo + __ldg(S.indices + io + prim * 3 + i));
        positions[i] = xyz(mul(proxy->worldTS, f4(v.pos.x, v.pos.y, v.pos.z, 1)));
        normals[i] = xyz(mul(proxy->worldTS, f4(v.nrm.x, v.nrm.y, v.nrm.z, 0)));
        lnrm[i] = v.nrm;
        ltan[i] = v.tan;
        tex[i] = v.uv;
        if (i == 0) matid = v.matSlot;
    }
    float3 e0 = positions[1] - positions[0], e1 = positions[2] - positions[0];
    float3 ray_cross_edge_1 = cross(ray_direction, e1);
    float rcp_det = 1.0f / dot(e0, ray_cross_edge_1);
    float3 ray_to_0 = ray_origin - positions[0];
    float by = rcp_det * dot(ray_to_0, ray_cross_edge_1);
    float3 edge_0_cross_0 = cross(e0, ray_to_0);
    float bz = -rcp_det * dot(ray_direction, edge_0_cross_0);
    float bx = 1.0f - (by + bz);
    PrimVertex r;
    r.pos = positions[0] * bx + (positions[1] * by + positions[2] * bz);
    r.nrm = normalize(normals[0] * bx + (normals[1] * by + normals[2] * bz));
    r.uv = tex[0] * bx + (tex[1] * by + tex[2] * bz);
    r.lnrm = lnrm[0] * bx + (lnrm[1] * by + lnrm[2] * bz);
    r.ltan = ltan[0] * bx + (ltan[1] * by + ltan[2] * bz);
    r.matSlot = matid;
    return r;
}

// The shading ray of a pixel goes through its corner (offset 0, Shading.slang:350-356); the visibility ray that stands in for the raster
// pass goes through its centre (offset 0.5), where the rasteriser samples coverage and depth (Rast.VisibilityPass.vert.slang:9-23).
GK_DEV float3 pixel_ray_dir(const GkUniformBufferObject* C, int x, int y, int w, int h, float offset = 0.0f) {
    float2 uv = make_float2(((float)x + offset) / (float)w * 2.0f - 1.0f, ((float)y + offset) / (float)h * 2.0f - 1.0f);
    float4 target = mul(C->ProjectionInverse, f4(uv.x, uv.y, 1, 1));
    float3 tn = normalize(xyz(target));
    float4 dir = mul(C->ModelViewInverse, f4(tn.x, tn.y, tn.z, 0));
    return normalize(xyz(dir));
}

// ------------------------------------------------------------------------------------------------ state machine
// GetRayColor up to the TraceRay call (Shading.slang:1005-1039)
GK_DEV void gen_closest(const SceneView& S, PathRegs& P, bool first) {
    GkMaterial mat = get_material(S, P.mat);
    const float startPosOffset = mat.MaterialModel == GK_MATERIAL_DIELECTRIC ? 0.0f : 1.0f;
    float roughness = mat.Fuzziness;
    const float dotValue = dot(P.dir, P.nrm);
    const bool BackFace = dotValue > 0;
    const float3 outwardNormal = BackFace ? -P.nrm : P.nrm;
    const float niOverNt = BackFace ? mat.RefractionIndex2 : (1 / mat.RefractionIndex2);
    const float cosine = dotValue > 0 ? mat.RefractionIndex * dotValue : -dotValue;
    const float reflectProb = Schlick(cosine, mat.RefractionIndex);
    const float metalProb = mat.Metalness;
    bool chanceReflect = (RandomFloat(P.rng) < reflectProb);
    bool chanceMetal = (RandomFloat(P.rng) < metalProb);
    bool chanceGGX = chanceReflect || chanceMetal;
    const float3 trace_next = chanceGGX ? reflect3(P.dir, outwardNormal) : outwardNormal;
    float3 trace_dir;
    if (chanceGGX) trace_dir = ggxSampling(P.rng, sqrtf(roughness), trace_next);
    else trace_dir = AlignWithNormal(RandomInHemiSphere1(P.rng), trace_next);
    if (first) {
        P.flags &= ~(FL_HIT_REFLECT | FL_HIT_METAL);
        if (chanceGGX) P.flags |= FL_HIT_REFLECT;
        if (chanceMetal) P.flags |= FL_HIT_METAL;
    }
    if (mat.MaterialModel == GK_MATERIAL_DIELECTRIC) {
        if (!chanceReflect) trace_dir = refract3(P.dir, outwardNormal, niOverNt);
    }
    P.dir = trace_dir;
    P.flags &= ~(FL_CHANCE_REFLECT | FL_BACKFACE_ND);
    if (chanceReflect) P.flags |= FL_CHANCE_REFLECT;
    if (BackFace && mat.MaterialModel != GK_MATERIAL_DIELECTRIC) P.flags |= FL_BACKFACE_ND;
    P.ro = P.pos + P.nrm * 0.001f * startPosOffset;  // HitNormalOffset, Core.PathTracing.comp.slang:18
    P.rd = trace_dir;
}

// FHardwareRayTracerV2::TraceRay after the query (Shading.slang:801-823) + the rest of GetRayColor (:1040-1081).
GK_DEV bool on_closest(const SceneView& S, PathRegs& P, const Hit& h) {
    if (h.inst != GK_NO_HIT) {
        const HitVertex hv = resolve_hit(S, P.ro, P.rd, h);
        P.mat = hv.mat;
        P.nrm = hv.nrm;
        P.pos = hv.pos;
        const float2 uv = hv.uv;
        GkMaterial hitMat = get_material(S, P.mat);
        float4 outAlbedo = f4(hitMat.Diffuse.x, hitMat.Diffuse.y, hitMat.Diffuse.z, hitMat.Diffuse.w);
        if (hitMat.DiffuseTextureId >= 0) outAlbedo = outAlbedo * sample_texture(S, (uint32_t)hitMat.DiffuseTextureId, uv);
        if (hitMat.MaterialModel == GK_MATERIAL_DIFFUSE_LIGHT || !(P.flags & FL_CHANCE_REFLECT)) P.rayColor = P.rayColor * outAlbedo;
        if (P.flags & FL_BACKFACE_ND) {
            P.rayColor = f4(0.5f, 0.5f, 0.5f, 0);
            return true;
        }
        if (hitMat.MaterialModel == GK_MATERIAL_DIFFUSE_LIGHT) return true;
        return false;
    }
    P.rayColor = P.rayColor * sky_color(S, P.rd);
    return true;
}

GK_DEV void set_stage(PathRegs& P, uint32_t st) { P.flags = (P.flags & ~FL_STAGE_MASK) | (st << FL_STAGE_SHIFT); }

// Shading.slang:1105-1117: reset the path to the primary hit for sample i (the first GetRayColor follows)
GK_DEV void begin_sample(const SceneView& S, const PathView& V, uint32_t pixel, PathRegs& P) {
    const float4 pa = V.prim_pos_mat[pixel], pb = V.prim_nrm_cnt[pixel];
    P.rayColor = f4(1, 1, 1, 1);
    float4 eye = mul(S.cam->ModelViewInverse, f4(0, 0, 0, 1));
    P.pos = f3(pa.x, pa.y, pa.z);
    P.dir = normalize(P.pos - xyz(eye));
    P.nrm = f3(pb.x, pb.y, pb.z);
    P.mat = __float_as_uint(pa.w);
    P.bounce = 0;
}
GK_DEV void start_sample(const SceneView& S, const PathView& V, uint32_t pixel, PathRegs& P) {
    begin_sample(S, V, pixel, P);
    set_stage(P, ST_FIRST);
    gen_closest(S, P, true);
}

struct PixelOut {
    RTView rt;
    float4 *radiance, *direct;
    size_t pixels;
};

GK_DEV void write_result(const PixelOut& O, const FrameCfg& cfg, const PathView& V, uint32_t pixel, float4 direct) {
    const float4 finalColor = V.final_color[pixel], finalRefl = V.final_refl[pixel];
    if (cfg.to_radiance) {
        O.radiance[pixel] = finalColor;
        O.radiance[O.pixels + pixel] = finalRefl;
        if (cfg.rank == 0) O.direct[pixel] = direct;
        return;
    }
    const uint32_t sampleTotal = __float_as_uint(V.prim_nrm_cnt[pixel].w) >> 16;
    float4 fc = finalColor / float(sampleTotal);   // Shading.slang:1175-1176
    float4 fr = finalRefl / float(sampleTotal);
    fc = fc + direct;                              // :1180
    store_h4(O.rt.single_diffuse, pixel, fc);
    store_h4(O.rt.single_specular, pixel, fr);
}

// Consumes the result of the pending ray and runs the pixel's program up to its next ray (or to completion).
// Written so that every heavy piece (hit resolve, direction sampling, sun cone, probe lookup, result write) has exactly one
// call site: the shade kernel's code must stay inside the instruction cache (ncu: 49 % "no instruction" stalls at 96 KB).
GK_DEV int advance(const SceneView& S, const PixelOut& O, const FrameCfg& cfg, const PathView& V, uint32_t pixel, PathRegs& P, const Hit& h) {
    enum { A_END_SAMPLE, A_AFTER_SUN, A_GEN_BOUNCE, A_SUN_RAY, A_DIRECT_RESULT, A_PROBE_END };
    const GkUniformBufferObject* C = S.cam;
    const uint32_t stage = (P.flags & FL_STAGE_MASK) >> FL_STAGE_SHIFT;
    const bool primDielectric = (P.flags & FL_PRIM_DIELECTRIC) != 0;
    const float3 lightVector = f3(C->SunDirection.x, C->SunDirection.y, C->SunDirection.z);
    const float4 sunColor = f4(C->SunColor.x, C->SunColor.y, C->SunColor.z, C->SunColor.w);
    int action = A_END_SAMPLE;
    float shadow = 0.0f;
    bool direct_stage = false;
    if (stage == ST_DIRECT) {  // FHardwareDirectIlluminatorV2::DirectIlluminate, :886-908
        shadow = (h.inst == GK_NO_HIT) ? 1.0f : 0.0f;
        action = A_DIRECT_RESULT;
    } else if (stage == ST_SUN) {  // :1139-1143
        if (h.inst == GK_NO_HIT) P.rayColor = P.rayColor * sunColor;
        else action = A_AFTER_SUN;
    } else {
        const bool exited = on_closest(S, P, h);
        if (stage == ST_FIRST) {
            if (!exited) {
                if (cfg.exit_after_first && !primDielectric) action = A_PROBE_END;  // :1120-1123
                else {
                    P.bounce = 1;
                    if (1u < P.maxBounces) action = A_GEN_BOUNCE;
                }
            }
        } else if (!exited) {  // :1130-1144
            if (C->HasSun && (RandomFloat(P.rng) < 0.5f)) action = A_SUN_RAY;
            else action = A_AFTER_SUN;
        }
    }
    if (action == A_AFTER_SUN) {  // :1146-1154
        const bool earlyExit = !primDielectric && (RandomFloat(P.rng) < 0.5f);  // ExitProbability
        if (P.bounce == P.maxBounces - 1u || earlyExit) action = A_PROBE_END;
        else {
            P.bounce++;
            action = A_GEN_BOUNCE;
        }
    }
    {   // path termination in the probe grid: every lane of the warp takes part in the lookups of the few that need one
        const bool probe_end = action == A_PROBE_END;
        const float4 gi = interpolateAmbientCubesCoop<FullAmbientCubeSampler>(S, LiveCubes{S.probes}, probe_end, P.pos, P.nrm);
        if (probe_end) {
            P.rayColor = P.rayColor * gi;
            action = A_END_SAMPLE;
        }
    }
    bool first = false;
    if (action == A_END_SAMPLE) {  // :1159-1171: fold the path into the pixel's running sums
        if (P.flags & FL_HIT_METAL) P.rayColor = P.rayColor * V.prim_albedo[pixel];
        float4* acc = (P.flags & FL_HIT_REFLECT) ? V.final_refl + pixel : V.final_color + pixel;
        *acc = *acc + P.rayColor;
        const uint32_t sampleEnd = __float_as_uint(V.prim_nrm_cnt[pixel].w) & 0xFFFFu;
        P.sample++;
        if (P.sample < sampleEnd) {
            begin_sample(S, V, pixel, P);
            first = true;
            action = A_GEN_BOUNCE;
        } else if (cfg.rank != 0) {  // all samples done, :1175-1183; only rank 0 owns the direct term
            shadow = 0.0f;
            action = A_DIRECT_RESULT;
            direct_stage = true;
        } else if (C->HasSun) {
            direct_stage = true;
            action = A_SUN_RAY;
        } else {
            shadow = 0.0f;
            action = A_DIRECT_RESULT;
        }
    }
    if (action == A_SUN_RAY) {  // cone-jittered shadow ray towards the sun: the bounce-loop NEE (:1137-1139) or DirectIlluminate (:897-899)
        const float3 cone = AlignWithNormal(RandomInCone(P.rng, GK_COS_QUARTER_DEGREE), lightVector);
        if (direct_stage) {
            const float4 pa = V.prim_pos_mat[pixel];
            P.ro = f3(pa.x, pa.y, pa.z);
            set_stage(P, ST_DIRECT);
        } else {
            P.ro = P.pos + P.nrm * 0.0f;  // TRACE_CORRECTION_OFFSET
            set_stage(P, ST_SUN);
        }
        P.rd = cone;
        return REQ_ANY;
    }
    if (action == A_DIRECT_RESULT) {
        const float4 pb = V.prim_nrm_cnt[pixel];
        const float d = fmaxf(dot(lightVector, normalize(f3(pb.x, pb.y, pb.z))), 0.0f) * GK_M_1_PI;
        const bool zero = cfg.rank != 0 && direct_stage;
        write_result(O, cfg, V, pixel, zero ? f4(0, 0, 0, 0) : sunColor * d * shadow);
        return REQ_DONE;
    }
    // A_GEN_BOUNCE: the next GetRayColor (first of a sample or a bounce), single call site
    set_stage(P, first ? ST_FIRST : ST_BOUNCE);
    gen_closest(S, P, first);
    return REQ_CLOSEST;
}

// ------------------------------------------------------------------------------------------------ state I/O
// Queue slot layout. A closest-hit request needs little of the path state back: the pending ray (its direction IS the path
// direction, and position / normal / material are overwritten by the hit or unused on a miss), the RNG, the throughput and
// two packed words, which ride in the .w lanes of the ray record (tmin / tmax of queue rays are constants):
//   slot_pixel 4 B | ray_o = (origin, flags) 16 B | ray_d = (direction, sample | bounce << 16) 16 B | rng 16 B | ray_color 16 B
// = 68 B per slot. A shadow ray (any-hit request) resumes the path at its current vertex, so only those slots also carry
// pos_mat / nrm / dir (48 B). This is what k_shade streams from and to HBM every wave.
GK_DEV void store_slot(const PathView& V, int qi, uint32_t slot, uint32_t pixel, int req, const PathRegs& P) {
    const size_t i = (size_t)qi * V.pixels + slot;
    V.slot_pixel[qi][slot] = pixel | (req == REQ_ANY ? 0x80000000u : 0u);
    V.ray_o[i] = f4(P.ro.x, P.ro.y, P.ro.z, __uint_as_float(P.flags));
    V.ray_d[i] = f4(P.rd.x, P.rd.y, P.rd.z, __uint_as_float(P.sample | (P.bounce << 16)));
    V.rng[i] = P.rng;
    V.ray_color[i] = P.rayColor;
    if (req == REQ_ANY) {
        V.pos_mat[i] = f4(P.pos.x, P.pos.y, P.pos.z, __uint_as_float(P.mat));
        V.nrm[i] = f4(P.nrm.x, P.nrm.y, P.nrm.z, 0.0f);
        V.dir[i] = f4(P.dir.x, P.dir.y, P.dir.z, 0.0f);
    }
}
// the four records every slot has, as loaded (k_shade issues these loads before it knows the queue size, see there)
#ifndef GK_SHADE_LDCS
#define GK_SHADE_LDCS 1  // slot and hit records are read once: ld.global.cs keeps them from displacing the shading records (k_shade 6.37 -> 6.33 ms)
#endif
struct SlotRaw { float4 o, d, ray_color; uint4 rng; };
GK_DEV SlotRaw load_slot_raw(const PathView& V, int qi, uint32_t slot) {
    const size_t i = (size_t)qi * V.pixels + slot;
    SlotRaw r;
#if GK_SHADE_LDCS
    // read once: streamed past the caches the gathers of the shading records live in
    r.o = __ldcs(V.ray_o + i); r.d = __ldcs(V.ray_d + i); r.rng = __ldcs(V.rng + i); r.ray_color = __ldcs(V.ray_color + i);
#else
    r.o = V.ray_o[i]; r.d = V.ray_d[i]; r.rng = V.rng[i]; r.ray_color = V.ray_color[i];
#endif
    return r;
}
GK_DEV void load_slot(const PathView& V, const SceneView& S, int qi, uint32_t slot, bool any, PathRegs& P, const SlotRaw* raw = nullptr) {
    const size_t i = (size_t)qi * V.pixels + slot;
    const SlotRaw r = raw ? *raw : load_slot_raw(V, qi, slot);
    const float4 o = r.o, d = r.d;
    P.ro = f3(o.x, o.y, o.z); P.flags = __float_as_uint(o.w);
    P.rd = f3(d.x, d.y, d.z);
    const uint32_t cnt = __float_as_uint(d.w);
    P.sample = cnt & 0xFFFF; P.bounce = cnt >> 16;
    P.rng = r.rng;
    P.rayColor = r.ray_color;
    if (any) {
        const float4 a = V.pos_mat[i], b = V.nrm[i], c = V.dir[i];
        P.pos = f3(a.x, a.y, a.z); P.mat = __float_as_uint(a.w);
        P.nrm = f3(b.x, b.y, b.z);
        P.dir = f3(c.x, c.y, c.z);
    } else {
        P.pos = f3(0, 0, 0); P.mat = 0u;
        P.nrm = f3(0, 0, 0);
        P.dir = P.rd;
    }
    P.maxBounces = (P.flags & FL_PRIM_DIELECTRIC) ? S.cam->MaxNumberOfBounces : S.cam->NumberOfBounces;
}

// Appends one entry to queue `qi` (one atomic per warp: ballot + popc + shuffle) and writes the slot.
GK_DEV void enqueue(const PathView& V, int qi, bool want, uint32_t pixel, int req, const PathRegs& P) {
    const unsigned mask = __ballot_sync(0xffffffffu, want);
    if (!mask) return;
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(mask) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(V.counts + qi, (uint32_t)__popc(mask));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (want) store_slot(V, qi, base + __popc(mask & ((1u << lane) - 1u)), pixel, req, P);
}

// ------------------------------------------------------------------------------------------------ kernels
// Primary visibility + G-buffer (FPathTracingRendererV2::Init/PrimaryHit, Shading.slang:1187-1241, with the visibility
// buffer lookup of FVisibilityBufferRayCasterV2::TracePrimaryRay, :338-494, replaced by a closest-hit cast through the pixel
// centre, where the raster pass samples), the buffer clear (Util.BufferClear.comp.slang:5-14), and the first ray of sample 0.
#ifndef GK_PRIMARY_MINB
#define GK_PRIMARY_MINB 6  // resident blocks per SM k_primary is compiled for: 80 registers + 128 B more spill instead of 122 (4 blocks): 0.740 -> 0.673 ms at 1080p
#endif
__global__ void __launch_bounds__(128, GK_PRIMARY_MINB) k_primary(SceneView S_, const __grid_constant__ GkUniformBufferObject cam, RTView R, const __grid_constant__ PathLanes L, FrameCfg cfg, unsigned long long* ray_counters) {
    const PathView& V = L.v[(blockIdx.x + blockIdx.y) % L.n];  // checkerboard of tiles over the lanes
    SceneView S = S_;
    S.cam = &cam;  // the camera block is read through the constant bank (uniform loads), not through L1
    const int x = blockIdx.x * 8 + (threadIdx.x & 7), y = blockIdx.y * 16 + (threadIdx.x >> 3);
    const bool inside = x < (int)R.w && y < (int)R.h;
    const uint32_t pixel = inside ? (uint32_t)y * R.w + (uint32_t)x : 0u;
    const GkUniformBufferObject* C = S.cam;
    PathRegs P;
    bool want = false;
    uint32_t nprimary = 0;
    if (inside) {
        const bool owner = true;  // every rank keeps its own (identical) G-buffer and MOTIONMOMENT in a multi-GPU frame
        P.rng = make_uint4((uint32_t)x, (uint32_t)y, cfg.frame_seed, 0u);  // InitRandomSeed, Const_Func.slang:237-240
        float4 origin4 = mul(C->ModelViewInverse, f4(0, 0, 0, 1));
        float3 origin = xyz(origin4);
        float3 ray_dir = pixel_ray_dir(C, x, y, R.w, R.h);
        Hit h;
        if (R.vis) {  // FVisibilityBufferRayCasterV2::TracePrimaryRay, Shading.slang:348-349: the visibility buffer the raster pass wrote
            const uint32_t v = R.vis[pixel];
            h.inst = v ? ((v >> 17) & 0x7FFFu) - 1u : GK_NO_HIT;
            h.prim = v & 0x1FFFFu;
        } else {
            trace_ray<false>(S.ts, origin, pixel_ray_dir(C, x, y, R.w, R.h, 0.5f), GK_EPS, GK_PRIMARY_T_MAX, h);
            nprimary++;
        }
        float4 zero = f4(0, 0, 0, 0);
        if (h.inst == GK_NO_HIT) {  // Shading.slang:1193-1201
            float4 sky = sky_color(S, ray_dir);
            if (cfg.to_radiance) {
                V.radiance[pixel] = zero;
                V.radiance[(size_t)R.w * R.h + pixel] = zero;
                if (cfg.rank == 0) V.direct[pixel] = sky;
            } else {
                store_h4(R.single_diffuse, pixel, sky);
                store_h4(R.single_specular, pixel, zero);
            }
            if (owner) {
                R.motion[pixel] = make_float2(0, 0);
                store_h4(R.albedo, pixel, f4(1, 1, 1, 1));
                store_h4(R.normal, pixel, f4(0, 1, 0, 1));
                R.objectid0[pixel] = GK_OBJECTID_MISS;
                R.primitiveid[pixel] = 0xFFFFFFFFu;
                R.prev_depth[pixel] = 100000.0f;
            }
        } else {
            PrimVertex iv = get_material_dataV2(S, h.inst, h.prim, origin, ray_dir);
            uint32_t node_idx = h.inst, prim = h.prim;
            float3 out_dir = ray_dir;
            PrimVertex fv = iv;
            // depth of field, Shading.slang:368-489
            float vertexDistance = length(iv.pos - origin);
            float cocRadius = 0.0f;
            if (fabsf(vertexDistance - C->FocusDistance) > 0.001f) cocRadius = (C->Aperture * fabsf(vertexDistance - C->FocusDistance)) / vertexDistance;
            if (cocRadius > 0.001f) {
                float2 diskSample = RandomInUnitDisk(P.rng);
                float3 right = normalize(cross(ray_dir, f3(0, 1, 0)));
                float3 circleEdge = iv.pos + right * cocRadius;
                float4 cp = mul(C->ViewProjection, f4(iv.pos.x, iv.pos.y, iv.pos.z, 1.0f));
                float4 ep = mul(C->ViewProjection, f4(circleEdge.x, circleEdge.y, circleEdge.z, 1.0f));
                float2 c2 = make_float2(cp.x / cp.w, cp.y / cp.w), e2 = make_float2(ep.x / ep.w, ep.y / ep.w);
                float screenSpaceRadius = length(e2 - c2);
                float2 pixelOffset = make_float2(diskSample.x * screenSpaceRadius * (float)R.w * 0.5f, diskSample.y * screenSpaceRadius * (float)R.h * 0.5f);
                int ox = min(max(x + (int)pixelOffset.x, 0), (int)R.w - 1), oy = min(max(y + (int)pixelOffset.y, 0), (int)R.h - 1);
                if (ox != x || oy != y) {
                    float3 fdir = pixel_ray_dir(C, ox, oy, R.w, R.h);
                    Hit h2;
                    if (R.vis) {
                        const uint32_t v = R.vis[(size_t)oy * R.w + ox];
                        h2.inst = v ? ((v >> 17) & 0x7FFFu) - 1u : GK_NO_HIT;
                        h2.prim = v & 0x1FFFFu;
                    } else {
                        trace_ray<false>(S.ts, origin, pixel_ray_dir(C, ox, oy, R.w, R.h, 0.5f), GK_EPS, GK_PRIMARY_T_MAX, h2);
                        nprimary++;
                    }
                    if (h2.inst != GK_NO_HIT) {
                        PrimVertex cand = get_material_dataV2(S, h2.inst, h2.prim, origin, fdir);
                        bool cancel = false;
                        if (fabsf(vertexDistance - C->FocusDistance) < C->FocusDistance * 0.1f) {
                            float depthDifference = vertexDistance - length(cand.pos - origin);
                            if (depthDifference > C->FocusDistance * 0.05f) cancel = true;
                        }
                        if (!cancel) { fv = cand; node_idx = h2.inst; prim = h2.prim; out_dir = fdir; }
                    }
                }
            }
            const GkNodeProxy* node = S.nodes + node_idx;
            float3 hpos = fv.pos;
            float3 hnrm = normalize(fv.nrm);
            uint32_t hmat = fetch_material_id(node, fv.matSlot);
            GkMaterial material = get_material(S, hmat);
            // FetchGBufferV2, Shading.slang:82-112
            float4 albedo = f4(material.Diffuse.x, material.Diffuse.y, material.Diffuse.z, material.Diffuse.w);
            if (material.DiffuseTextureId >= 0) albedo = albedo * sample_texture(S, (uint32_t)material.DiffuseTextureId, fv.uv);
            float roughness = material.Fuzziness;
            if (material.MRATextureId >= 0) roughness = roughness * sample_texture(S, (uint32_t)material.MRATextureId, fv.uv).y;
            if (material.NormalTextureId >= 0) {  // mikkTSpace in the local frame, then to world space (:74-80, 100-107)
                const float4 t = sample_texture(S, (uint32_t)material.NormalTextureId, fv.uv);
                const float3 vNt = f3(t.x, t.y, t.z) * 2.0f - f3(1.0f, 1.0f, 1.0f);
                const float3 tg = f3(fv.ltan.x, fv.ltan.y, fv.ltan.z);
                const float3 vB = cross(fv.lnrm, tg) * fv.ltan.w;
                const float3 nl = normalize(tg * vNt.x + vB * vNt.y + fv.lnrm * vNt.z);
                hnrm = normalize(xyz(mul(node->worldTS, f4(nl.x, nl.y, nl.z, 0.0f))));
            }
            uint16_t moment = R.moment[pixel];
            {   // CalculateMotionVectorV2, Shading.slang:50-72
                float4 Pw = f4(hpos.x, hpos.y, hpos.z, 1);
                float4 cur = mul(C->ViewProjectionUnJit, Pw);
                float2 currfpos = make_float2(cur.x / cur.w * 0.5f, cur.y / cur.w * 0.5f);
                float4 prevPos = mul(node->combinedPrevTS, Pw);
                // mul(mul(PrevVP, combinedPrevTS), P): the matrix product first, like the shader
                GkMat4 pm;
                for (int c = 0; c < 4; ++c)
                    for (int r = 0; r < 4; ++r)
                        pm.m[c][r] = C->PrevViewProjectionUnJit.m[0][r] * node->combinedPrevTS.m[c][0] + C->PrevViewProjectionUnJit.m[1][r] * node->combinedPrevTS.m[c][1] +
                                     C->PrevViewProjectionUnJit.m[2][r] * node->combinedPrevTS.m[c][2] + C->PrevViewProjectionUnJit.m[3][r] * node->combinedPrevTS.m[c][3];
                float4 prev = mul(pm, Pw);
                float2 prevfpos = make_float2(prev.x / prev.w * 0.5f, prev.y / prev.w * 0.5f);
                if (length(hpos - xyz(prevPos)) > 0.02f) moment = (uint16_t)C->TemporalFrames;
                else moment = moment > 0 ? moment - 1 : 0;
                float2 motion = prevfpos - currfpos;
                if (owner) {
                    R.moment[pixel] = moment;
                    R.motion[pixel] = make_float2(motion.x * (float)R.w, motion.y * (float)R.h);
                }
            }
            if (owner) {
                float4 clip = mul(C->ViewProjection, f4(hpos.x, hpos.y, hpos.z, 1.0f));
                store_h4(R.albedo, pixel, albedo);
                store_h4(R.normal, pixel, f4(hnrm.x, hnrm.y, hnrm.z, roughness));
                R.objectid0[pixel] = node->instanceId;
                R.primitiveid[pixel] = prim;
                R.prev_depth[pixel] = clip.z / clip.w;
            }
            if (material.MaterialModel != GK_MATERIAL_DIELECTRIC) hpos = hpos - out_dir * 0.03f;  // :1217-1220
            if (material.MaterialModel == GK_MATERIAL_DIFFUSE_LIGHT) {  // :1093-1099
                float4 e = f4(material.Diffuse.x, material.Diffuse.y, material.Diffuse.z, material.Diffuse.w);
                if (cfg.to_radiance) {
                    V.radiance[pixel] = zero;
                    V.radiance[(size_t)R.w * R.h + pixel] = zero;
                    if (cfg.rank == 0) V.direct[pixel] = e;
                } else {
                    store_h4(R.single_diffuse, pixel, e);
                    store_h4(R.single_specular, pixel, zero);
                }
            } else {
                int sampleMultiplier = moment > 0 ? 4 : 1;  // Core.PathTracing.comp.slang:28
                int sample = max(1, (int)C->NumberOfSamples * sampleMultiplier);  // SampleDownscale = 1
                uint32_t s0 = (uint32_t)((uint64_t)cfg.rank * sample / cfg.world), s1 = (uint32_t)((uint64_t)(cfg.rank + 1) * sample / cfg.world);
                V.prim_pos_mat[pixel] = f4(hpos.x, hpos.y, hpos.z, __uint_as_float(hmat));
                V.prim_nrm_cnt[pixel] = f4(hnrm.x, hnrm.y, hnrm.z, __uint_as_float(s1 | ((uint32_t)sample << 16)));
                V.prim_albedo[pixel] = albedo;
                V.final_color[pixel] = zero;
                V.final_refl[pixel] = zero;
                P.flags = material.MaterialModel == GK_MATERIAL_DIELECTRIC ? FL_PRIM_DIELECTRIC : 0u;
                P.maxBounces = (P.flags & FL_PRIM_DIELECTRIC) ? C->MaxNumberOfBounces : C->NumberOfBounces;
                P.sample = s0;
                if (s0 < s1) {
                    start_sample(S, V, pixel, P);
                    want = true;
                } else {  // this rank has no samples for the pixel (spp < world)
                    if (cfg.to_radiance) {
                        V.radiance[pixel] = zero;
                        V.radiance[(size_t)R.w * R.h + pixel] = zero;
                        if (cfg.rank == 0) V.direct[pixel] = zero;
                    }
                }
            }
        }
    }
    enqueue(V, 0, want, pixel, REQ_CLOSEST, P);
    // ray statistics: one atomic per warp
    for (int o = 16; o > 0; o >>= 1) nprimary += __shfl_xor_sync(0xffffffffu, nprimary, o);
    if ((threadIdx.x & 31) == 0 && nprimary) atomicAdd(ray_counters + 0, (unsigned long long)nprimary);
}

// Device-side end of the wavefront: counts[3] = (queue index + 1) of the queue the wavefront stopped at, 0 while it runs. The first
// k_trace that finds its input queue at or below `stop_below` raises it; every kernel of the loop launched after that returns at once
// and k_tail picks the queue up. All blocks of a kernel take the same decision: the flag only ever goes from 0 to non-zero, and the
// queue size a k_trace reads is not written while it runs.
GK_DEV bool wave_stopped(const PathView& V, int qi, uint32_t n, uint32_t stop_below) {
    const bool stopped_before = V.counts[3] != 0;
    const bool lead = blockIdx.x == 0 && threadIdx.x == 0;
    // the host learns the queue sizes from these stores (mapped pinned memory, no copy engine); 0 once the wavefront has stopped
    if (lead && V.publish) *V.publish = stopped_before ? 0u : n;
    if (!stopped_before && n > stop_below) return false;
    if (lead && !stopped_before) V.counts[3] = (uint32_t)qi + 1u;
    return true;
}

// Ray reordering (GKCU_SORT_RAYS): key = Morton code of the origin's cell on a 2^bits grid over the TLAS root box, then 6 bits of direction
// (octant + dominant axis + a half-space bit). The rays stay where they are; k_trace walks them through the sorted index, so hits land in
// the slots the shade kernel expects and the result cannot depend on the order.
GK_DEV unsigned long long spread3(uint32_t v) {  // 21 bits -> every third bit
    unsigned long long x = v & 0x1FFFFFull;
    x = (x | x << 32) & 0x1F00000000FFFFull;
    x = (x | x << 16) & 0x1F0000FF0000FFull;
    x = (x | x << 8) & 0x100F00F00F00F00Full;
    x = (x | x << 4) & 0x10C30C30C30C30C3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}
__global__ void k_ray_keys(PathView V, int qi, size_t half, uint32_t nmax, float3 lo, float3 inv_extent, int bits, unsigned long long* __restrict__ keys, uint32_t* __restrict__ vals) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nmax) return;
    unsigned long long key = ~0ull;
    if (i < V.counts[qi]) {
        const float4 o = V.ray_o[(size_t)qi * half + i], d = V.ray_d[(size_t)qi * half + i];
        const float scale = (float)(1u << bits);
        const uint32_t cx = (uint32_t)fminf(fmaxf((o.x - lo.x) * inv_extent.x * scale, 0.0f), scale - 1.0f);
        const uint32_t cy = (uint32_t)fminf(fmaxf((o.y - lo.y) * inv_extent.y * scale, 0.0f), scale - 1.0f);
        const uint32_t cz = (uint32_t)fminf(fmaxf((o.z - lo.z) * inv_extent.z * scale, 0.0f), scale - 1.0f);
        const unsigned long long m = spread3(cx) | spread3(cy) << 1 | spread3(cz) << 2;
        const float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
        const uint32_t major = ax >= ay && ax >= az ? 0u : (ay >= az ? 1u : 2u);
        const uint32_t oct = (d.x < 0.0f ? 1u : 0u) | (d.y < 0.0f ? 2u : 0u) | (d.z < 0.0f ? 4u : 0u);
        key = m << 5 | oct << 2 | major;
    }
    keys[i] = key;
    vals[i] = i;
}

// Persistent-thread traversal: every lane owns one ray, and lanes whose ray has terminated pull the next ray from the queue
// (one atomic per warp) as soon as the number of busy lanes drops below `refill_below`. VARIANT 5 (default): each iteration
// the warp votes and executes only the step kind most lanes are waiting for (node walk, instance entry or triangle leaf);
// lanes in another state sit the iteration out instead of serialising a second code path; compiled for GK_TRACE_MINB
// blocks/SM with the world-space ray parked in shared memory. VARIANT 0: every lane advances one step of whatever kind
// per iteration (the baseline the vote is measured against, profiles/r01_experiments.md).
template <int VARIANT, int FMT>
__global__ void __launch_bounds__(128, VARIANT == 5 ? GK_TRACE_MINB : 1) k_trace(TraceScene ts, PathView V, int qi, size_t half, int refill_below, uint32_t stop_below) {
    const uint32_t n = V.counts[qi];
    if (wave_stopped(V, qi, n, stop_below)) return;
    // the other queue (this wave's output, last wave's input) is idle while this kernel runs: clear its size here instead of
    // with a memset between the kernels
    if (blockIdx.x == 0 && threadIdx.x == 0) V.counts[qi ^ 1] = 0;
    const int lane = threadIdx.x & 31;
    const float4* ray_o = V.ray_o + (size_t)qi * half;
    const float4* ray_d = V.ray_d + (size_t)qi * half;
    const uint32_t* slot_pixel = V.slot_pixel[qi];
    __shared__ float s_world[VARIANT == 5 ? 16 * 128 : 1];
    typename std::conditional<VARIANT == 5, WorldSmemX<128>, WorldRegs>::type world;
    if constexpr (VARIANT == 5) world.p = s_world + threadIdx.x;
    typename std::conditional<GK_TOS_REG && VARIANT == 5 && FMT == 4, TosStack, LocalStack>::type stack;
    Trav T;
    T.done = true;
    uint32_t slot = 0;
    bool busy = false;
    bool drained = false;  // the queue has no more rays
    for (;;) {
        if (!drained) {
            const unsigned idle = __ballot_sync(0xffffffffu, !busy);
            if (idle) {
                const int leader = __ffs(idle) - 1;
                uint32_t base = 0;
                if (lane == leader) base = atomicAdd(V.counts + 2, (uint32_t)__popc(idle));
                base = __shfl_sync(0xffffffffu, base, leader);
                if (!busy) {
                    slot = base + __popc(idle & ((1u << lane) - 1u));
                    if (slot < n) {
                        if (V.order) slot = __ldcs(V.order + slot);
                        const float4 o = __ldcs(ray_o + slot), d = __ldcs(ray_d + slot);  // streamed: keep L1 for nodes and stacks
                        trav_init<FMT>(T, world, ts, f3(o.x, o.y, o.z), f3(d.x, d.y, d.z), GK_EPS, GK_PT_MAX_TRACE_DISTANCE, (__ldcs(slot_pixel + slot) & 0x80000000u) != 0);
                        busy = true;
                    }
                }
                if (base + (uint32_t)__popc(idle) >= n) drained = true;
            }
        }
        if (!__any_sync(0xffffffffu, busy)) return;
        if (VARIANT == 0) {
            do {
                if (busy && !T.done) {
                    if (!(T.cur & GK_LEAF_BIT)) trav_node_step<FMT>(T, world, ts, stack);
                    else trav_leaf_step<FMT>(T, world, ts, stack);
                }
                if (busy && T.done) {
                    __stcs(V.hit_tuv + slot, f4(T.best.t, T.best.u, T.best.v, __uint_as_float(T.best.inst)));
                    __stcs(V.hit_prim + slot, T.best.prim);
                    busy = false;
                }
            } while (__popc(__ballot_sync(0xffffffffu, busy)) >= (drained ? 1 : refill_below));
        } else {
            for (;;) {
                const bool live = busy && !T.done;
                const bool leaf = (T.cur & GK_LEAF_BIT) != 0;
                const bool wN = live && !leaf;
                const bool wI = live && leaf && T.cur_inst == GK_NO_HIT;
                const bool wT = live && leaf && T.cur_inst != GK_NO_HIT;
                const unsigned bN = __ballot_sync(0xffffffffu, wN), bI = __ballot_sync(0xffffffffu, wI), bT = __ballot_sync(0xffffffffu, wT);
                const int nN = __popc(bN), nI = __popc(bI), nT = __popc(bT);
                if (nN + nI + nT < (drained ? 1 : refill_below)) break;  // too few lanes left: refill (or finish)
                if (nN >= nT && nN >= nI) {
                    if (wN) trav_node_step<FMT>(T, world, ts, stack);
#if GK_NODE_REPEAT > 1
                    // most lanes continue with another node: take a second node step before the next vote
                    if (wN && !T.done && !(T.cur & GK_LEAF_BIT)) trav_node_step<FMT>(T, world, ts, stack);
#endif
                } else if (nT >= nI) {
                    if (wT) trav_leaf_step<FMT>(T, world, ts, stack);
                } else {
                    if (wI) trav_leaf_step<FMT>(T, world, ts, stack);
                }
                if (busy && T.done) {
                    __stcs(V.hit_tuv + slot, f4(T.best.t, T.best.u, T.best.v, __uint_as_float(T.best.inst)));
                    __stcs(V.hit_prim + slot, T.best.prim);
                    busy = false;
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------- pool tracer
// k_trace_pool: the rays of a block live in a POOL of slots in shared memory instead of in the registers of a fixed lane, and
// every slot waits in the queue of the step kind it needs next (node walk, triangle leaf, instance entry, stack pop, refill).
// A phase gives each warp up to 32 slots of ONE kind - full chunks first - so the step code runs with (almost) all lanes active,
// where the lane-owned loop of k_trace executes with ~15 of 32 (ncu, profiles/). The price is the state round trip through shared
// memory (about 20 words per step), one block barrier per phase and a traversal stack in global memory ([depth][slot] per block,
// L1-resident like the local-memory stack it replaces). Per ray the sequence of node tests, pushes, pops and triangle tests is
// exactly that of trace_ray, so the result is bit-identical. Node4 only.
#ifndef GK_POOL_MINB
#define GK_POOL_MINB 4
#endif
#define GK_POOL 256
#ifndef GK_POOL_MIN_CHUNK
#define GK_POOL_MIN_CHUNK 20  // a rest smaller than this waits for more slots of its kind
#endif
#ifndef GK_POOL_ROUNDS
#define GK_POOL_ROUNDS 2      // chunks per warp and phase
#endif
enum { PK_FREE = 0, PK_POP = 1, PK_TRI = 2, PK_INST = 3, PK_NODE = 4, PK_COUNT = 5 };
#define PP_SGN_SHIFT 16
#define PP_ANY (1u << 19)
#define PP_HAS_HIT (1u << 20)

struct PoolSmem {
    float ro[3][GK_POOL], idir[3][GK_POOL];  // current-space ray
    float world[9][GK_POOL];                 // world-space origin, direction, reciprocal direction (WorldSmem layout)
    float sh[3][GK_POOL];                    // watertight-test shear of the current space
    float best_t[GK_POOL];
    uint32_t shk[GK_POOL];                   // kx | ky << 2 | kz << 4
    uint32_t cur[GK_POOL], nbase[GK_POOL], tbase[GK_POOL], cinst[GK_POOL], pack[GK_POOL], qslot[GK_POOL];
    uint16_t ring[PK_COUNT][2 * GK_POOL];    // per kind: ring buffer of waiting slots; twice the pool, so that the slots appended in a phase
                                             // (at most the pool) can never wrap onto entries dealt in the same phase and not read yet
    uint32_t pushed[3][8];                   // slots appended per kind in phase p, p + 1, p + 2 (rotating); [7]: the ray queue ran empty
};

// traversal stack of one slot in global memory: entry d of slot s of this block at [d][s]
struct PoolStack {
    uint2* base;
    GK_DEV void push(int& sp, uint2 v) { base[(size_t)sp * GK_POOL] = v; ++sp; }
    GK_DEV uint2 pop(int& sp) { --sp; return base[(size_t)sp * GK_POOL]; }
    GK_DEV void put(int at, uint2 v) { base[(size_t)at * GK_POOL] = v; }
};

template <int FMT>
__global__ void __launch_bounds__(128, GK_POOL_MINB) k_trace_pool(TraceScene ts, PathView V, int qi, size_t half, uint2* __restrict__ stack_mem, uint32_t stop_below) {
    static_assert(FMT == 4, "the pool tracer walks Node4 trees");
    __shared__ PoolSmem sm;
    const uint32_t n = V.counts[qi];
    if (wave_stopped(V, qi, n, stop_below)) return;
    if (blockIdx.x == 0 && threadIdx.x == 0) V.counts[qi ^ 1] = 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float4* ray_o = V.ray_o + (size_t)qi * half;
    const float4* ray_d = V.ray_d + (size_t)qi * half;
    const uint32_t* slot_pixel = V.slot_pixel[qi];
    uint2* const my_stack = stack_mem + (size_t)blockIdx.x * GK_TRAV_STACK * GK_POOL;
    for (int i = threadIdx.x; i < GK_POOL; i += blockDim.x) sm.ring[PK_FREE][i] = (uint16_t)i;
    if (threadIdx.x < 24) (&sm.pushed[0][0])[threadIdx.x] = 0;
    bool drained = false;  // uniform: read from the counter set of the previous phase
    uint32_t head[PK_COUNT], tail[PK_COUNT];
#pragma unroll
    for (int k = 0; k < PK_COUNT; ++k) { head[k] = 0; tail[k] = 0; }
    tail[PK_FREE] = GK_POOL;
    uint32_t p_prev = 2, p_cur = 0, p_next = 1;  // counter sets: read, push, clear
    __syncthreads();
    for (uint32_t phase = 0;; ++phase) {
        if (phase) {
#pragma unroll
            for (int k = 0; k < PK_COUNT; ++k) tail[k] += sm.pushed[p_prev][k];
            drained = drained || sm.pushed[p_prev][7] != 0;
        }
        if (threadIdx.x < 8) sm.pushed[p_next][threadIdx.x] = 0;
        if (drained && tail[PK_FREE] - head[PK_FREE] == GK_POOL) break;  // every slot is free and no ray is left
        // ---- deal chunks of up to 32 waiting slots to the warps (every thread computes the same deal from the same counters): a kind
        // offers its full chunks and a rest of at least GK_POOL_MIN_CHUNK slots; smaller rests keep waiting for company unless nothing
        // else is left, then the largest one runs. Up to GK_POOL_ROUNDS chunks per warp and phase, chunk c goes to warp c % 4.
        uint32_t avail[PK_COUNT], nch[PK_COUNT], pre[PK_COUNT + 1];
#pragma unroll
        for (int k = 0; k < PK_COUNT; ++k) avail[k] = tail[k] - head[k];
        if (drained) avail[PK_FREE] = 0;
        pre[0] = 0;
#pragma unroll
        for (int k = 0; k < PK_COUNT; ++k) {
            nch[k] = (avail[k] >> 5) + ((avail[k] & 31u) >= (uint32_t)GK_POOL_MIN_CHUNK ? 1u : 0u);
            pre[k + 1] = pre[k] + nch[k];
        }
        if (pre[PK_COUNT] == 0) {  // only small rests are left: run the largest
            uint32_t best = 0;
            int kb = -1;
#pragma unroll
            for (int k = 0; k < PK_COUNT; ++k)
                if (avail[k] > best) { best = avail[k]; kb = k; }
#pragma unroll
            for (int k = 0; k < PK_COUNT; ++k) {
                nch[k] = k == kb ? 1u : 0u;
                pre[k + 1] = pre[k] + nch[k];
            }
        }
        const uint32_t total = min(pre[PK_COUNT], 4u * GK_POOL_ROUNDS);
        uint32_t begin_of[PK_COUNT];
#pragma unroll
        for (int k = 0; k < PK_COUNT; ++k) {
            begin_of[k] = head[k];
            const uint32_t dealt = min(nch[k], total > pre[k] ? total - pre[k] : 0u);
            head[k] += min(avail[k], dealt << 5);
        }
      for (uint32_t round = 0; round < (uint32_t)GK_POOL_ROUNDS; ++round) {
        const uint32_t chunk = (uint32_t)warp + 4u * round;
        if (chunk >= total) break;  // warp-uniform
        int my_kind = 0;
        uint32_t my_begin = 0, my_cnt = 0;
#pragma unroll
        for (int k = 0; k < PK_COUNT; ++k) {
            if (chunk >= pre[k] && chunk < pre[k + 1]) {
                my_kind = k;
                my_begin = begin_of[k] + ((chunk - pre[k]) << 5);
                my_cnt = min(32u, avail[k] - ((chunk - pre[k]) << 5));
            }
        }
        // ---- run one step of my_kind on my chunk
        const bool act = (uint32_t)lane < my_cnt;
        const uint32_t slot = act ? sm.ring[my_kind][(my_begin + lane) & (2 * GK_POOL - 1)] : 0u;
        int next = -1;  // queue the slot goes to afterwards
        if (my_kind == PK_FREE) {
            // refill: one atomic per warp takes the next rays from the queue
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(V.counts + 2, my_cnt);
            base = __shfl_sync(0xffffffffu, base, 0);
            if (base + my_cnt >= n && lane == 0) sm.pushed[p_cur][7] = 1;
            if (act) {
                const uint32_t q = base + lane;
                next = PK_FREE;
                if (q < n) {
                    const float4 o = __ldcs(ray_o + q), d = __ldcs(ray_d + q);
                    const bool any = (__ldcs(slot_pixel + q) & 0x80000000u) != 0;
                    Trav T;
                    WorldSmem<GK_POOL> world;
                    world.p = &sm.world[0][slot];
                    trav_init<FMT>(T, world, ts, f3(o.x, o.y, o.z), f3(d.x, d.y, d.z), GK_EPS, GK_PT_MAX_TRACE_DISTANCE, any);
                    // the miss record; a hit overwrites it
                    __stcs(V.hit_tuv + q, f4(T.best.t, 0.0f, 0.0f, __uint_as_float(GK_NO_HIT)));
                    __stcs(V.hit_prim + q, 0u);
                    if (!T.done) {
                        sm.ro[0][slot] = T.ro.x; sm.ro[1][slot] = T.ro.y; sm.ro[2][slot] = T.ro.z;
                        sm.idir[0][slot] = T.idir.x; sm.idir[1][slot] = T.idir.y; sm.idir[2][slot] = T.idir.z;
                        sm.best_t[slot] = T.best.t;
                        sm.cur[slot] = 0u;
                        sm.nbase[slot] = 0xFFFFFFFFu;
                        sm.tbase[slot] = 0u;
                        sm.cinst[slot] = GK_NO_HIT;
                        sm.pack[slot] = (T.sgn << PP_SGN_SHIFT) | (any ? PP_ANY : 0u);
                        sm.qslot[slot] = q;
                        next = PK_NODE;
                    }
                }
            }
        } else if (my_kind == PK_NODE) {
            if (act) {
                Trav T;
                const uint32_t pk = sm.pack[slot], nb = sm.nbase[slot];
                set_space<FMT>(T, f3(sm.ro[0][slot], sm.ro[1][slot], sm.ro[2][slot]), f3(sm.idir[0][slot], sm.idir[1][slot], sm.idir[2][slot]));
                T.tmin = GK_EPS;
                T.best.t = sm.best_t[slot];
                T.cur = sm.cur[slot];
                T.cur_inst = 0;  // only feeds the statistics build
                T.sp = (int)(pk & 0xFFu);
                T.nodes = reinterpret_cast<const uint4*>(nb == 0xFFFFFFFFu ? ts.tlas : ts.blas + nb);
                PoolStack stack{my_stack + slot};
                if (trav_node_core4(T, stack)) {
                    sm.cur[slot] = T.cur;
                    next = !(T.cur & GK_LEAF_BIT) ? PK_NODE : (nb == 0xFFFFFFFFu ? PK_INST : PK_TRI);
                } else {
                    next = PK_POP;
                }
                sm.pack[slot] = (pk & ~0xFFu) | (uint32_t)T.sp;
            }
        } else if (my_kind == PK_INST || my_kind == PK_TRI || my_kind == PK_POP) {
            if (act) {
                Trav T;
                uint32_t pk = sm.pack[slot];
                WorldSmem<GK_POOL> world;
                world.p = &sm.world[0][slot];
                T.cur = sm.cur[slot];
                T.cur_inst = sm.cinst[slot];
                T.sp = (int)(pk & 0xFFu);
                T.sp_inst = (int)((pk >> 8) & 0xFFu);
                T.best.t = sm.best_t[slot];
                T.done = false;
                bool pop = true, space_changed = false;
                if (my_kind == PK_INST) {
                    if (trav_enter_instance<FMT>(T, world, ts, T.cur & 0x0FFFFFFFu)) {
                        const InstanceRec* ir = ts.inst + T.cur_inst;
                        sm.nbase[slot] = __ldg(&ir->node_base);
                        sm.tbase[slot] = __ldg(&ir->tri_base);
                        sm.sh[0][slot] = T.sh.Sx; sm.sh[1][slot] = T.sh.Sy; sm.sh[2][slot] = T.sh.Sz;
                        sm.shk[slot] = (uint32_t)T.sh.kx | ((uint32_t)T.sh.ky << 2) | ((uint32_t)T.sh.kz << 4);
                        pop = false;
                        space_changed = true;
                        next = PK_NODE;
                    }
                } else if (my_kind == PK_TRI) {
                    const float3 ro = f3(sm.ro[0][slot], sm.ro[1][slot], sm.ro[2][slot]);
                    RayShear sh;
                    const uint32_t k3 = sm.shk[slot];
                    sh.kx = (int)(k3 & 3u); sh.ky = (int)((k3 >> 2) & 3u); sh.kz = (int)((k3 >> 4) & 3u);
                    sh.Sx = sm.sh[0][slot]; sh.Sy = sm.sh[1][slot]; sh.Sz = sm.sh[2][slot];
                    const float4* tris = ts.tris + (size_t)sm.tbase[slot] * 3;
                    const uint32_t q = sm.qslot[slot];
                    const uint32_t first = T.cur & 0x0FFFFFFFu, count = ((T.cur >> 28) & 7u) + 1u;
                    for (uint32_t i = 0; i < count; ++i) {
                        const float4* tp = tris + (size_t)(first + i) * 3;
                        const float4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2);
                        float t, b1, b2;
                        if (!tri_test(f3(a.x, a.y, a.z), f3(b.x, b.y, b.z), f3(c.x, c.y, c.z), ro, sh, t, b1, b2)) continue;
                        if (!(t > GK_EPS && t < GK_PT_MAX_TRACE_DISTANCE)) continue;
                        const uint32_t p = __float_as_uint(a.w);
                        if (pk & PP_ANY) {  // first hit ends the query
                            __stcs(V.hit_tuv + q, f4(t, 0.0f, 0.0f, __uint_as_float(T.cur_inst)));
                            __stcs(V.hit_prim + q, p);
                            T.done = true;
                            break;
                        }
                        bool better = !(pk & PP_HAS_HIT) || t < T.best.t;
                        if (!better && t == T.best.t) {  // tie: (instance, primitive) order; the record of the best hit is in global memory
                            const uint32_t bi = __float_as_uint(__ldcg(&V.hit_tuv[q].w)), bp = __ldcg(V.hit_prim + q);
                            better = T.cur_inst < bi || (T.cur_inst == bi && p < bp);
                        }
                        if (better) {
                            T.best.t = t;
                            pk |= PP_HAS_HIT;
                            __stcg(V.hit_tuv + q, f4(t, b1, b2, __uint_as_float(T.cur_inst)));
                            __stcg(V.hit_prim + q, p);
                        }
                    }
                    sm.best_t[slot] = T.best.t;
                    if (T.done) { pop = false; next = PK_FREE; }
                }
                if (pop) {
                    const uint32_t inst_before = T.cur_inst;
                    PoolStack stack{my_stack + slot};
                    trav_pop<FMT>(T, world, ts, stack);
                    if (T.done) next = PK_FREE;
                    else next = !(T.cur & GK_LEAF_BIT) ? PK_NODE : (T.cur_inst == GK_NO_HIT ? PK_INST : PK_TRI);
                    space_changed = T.cur_inst != inst_before;  // left the instance: back in world space
                    if (space_changed) sm.nbase[slot] = 0xFFFFFFFFu;
                    sm.cur[slot] = T.cur;
                }
                if (space_changed) {
                    sm.ro[0][slot] = T.ro.x; sm.ro[1][slot] = T.ro.y; sm.ro[2][slot] = T.ro.z;
                    sm.idir[0][slot] = T.idir.x; sm.idir[1][slot] = T.idir.y; sm.idir[2][slot] = T.idir.z;
                    sm.cinst[slot] = T.cur_inst;
                    pk = (pk & ~(7u << PP_SGN_SHIFT)) | (T.sgn << PP_SGN_SHIFT);
                    if (my_kind == PK_INST) sm.cur[slot] = T.cur;
                }
                sm.pack[slot] = (pk & ~0xFFFFu) | (uint32_t)T.sp | ((uint32_t)T.sp_inst << 8);
            }
        }
        // ---- append the slots to the queues of their next step (one shared-memory atomic per kind and warp)
        {
            const unsigned live = __ballot_sync(0xffffffffu, next >= 0);
            if (next >= 0) {
                const unsigned peers = __match_any_sync(live, next);
                const int leader = __ffs(peers) - 1;
                uint32_t base = 0;
                if (lane == leader) base = atomicAdd(&sm.pushed[p_cur][next], (uint32_t)__popc(peers));
                base = __shfl_sync(peers, base, leader);
                uint32_t t0 = 0;
#pragma unroll
                for (int k = 0; k < PK_COUNT; ++k)
                    if (k == next) t0 = tail[k];
                sm.ring[next][(t0 + base + __popc(peers & ((1u << lane) - 1u))) & (2 * GK_POOL - 1)] = (uint16_t)slot;
            }
        }
      }
        const uint32_t tp = p_prev; p_prev = p_cur; p_cur = p_next; p_next = tp;
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------------- warp-pool tracer
// k_trace_wpool: the pool idea without the block. Every WARP owns GK_WPOOL ray slots in shared memory and one queue per step kind;
// an iteration takes up to 32 slots of the kind that has a full chunk waiting (else the largest rest) and runs that step with
// (almost) all lanes active. No block barrier, no atomics - the queues are private to the warp - and a smaller state record: the
// world-space ray is re-read from the ray queue when an instance is entered or left, the BLAS bases from the instance record.
// Per ray the sequence of node tests, pushes, pops and triangle tests is that of trace_ray: bit-identical results. Node4 only.
#ifndef GK_WPOOL
#define GK_WPOOL 128
#endif
#ifndef GK_WPOOL_MINB
#define GK_WPOOL_MINB 4
#endif
static_assert(GK_WPOOL <= 128 && GK_WPOOL % 32 == 0, "slot ids are bytes, rings have 128 entries");

struct WPoolSmem {
    float ro[3][GK_WPOOL], idir[3][GK_WPOOL];  // current-space ray
    float sh[3][GK_WPOOL];                     // watertight-test shear of the current space
    float best_t[GK_WPOOL];
    uint32_t shk[GK_WPOOL];                    // kx | ky << 2 | kz << 4
    uint32_t cur[GK_WPOOL], nbase[GK_WPOOL], cinst[GK_WPOOL], pack[GK_WPOOL], qslot[GK_WPOOL];
    uint8_t ring[PK_COUNT][128];               // per kind: ring buffer of waiting slots
};

// the world-space ray of a slot, re-read from the ray queue (instance entry / exit only)
struct WorldQueue {
    static constexpr bool kKeepsTlasSpace = false;
    const float4 *ro, *rd;
    GK_DEV float3 o() const { const float4 v = __ldg(ro); return f3(v.x, v.y, v.z); }
    GK_DEV float3 d() const { const float4 v = __ldg(rd); return f3(v.x, v.y, v.z); }
    GK_DEV float3 widir() const { const float3 v = d(); return f3(safe_rcp(v.x), safe_rcp(v.y), safe_rcp(v.z)); }
    GK_DEV void set(float3, float3, float3) {}
};
struct WPoolStack {
    uint2* base;
    GK_DEV void push(int& sp, uint2 v) { base[(size_t)sp * GK_WPOOL] = v; ++sp; }
    GK_DEV uint2 pop(int& sp) { --sp; return base[(size_t)sp * GK_WPOOL]; }
    GK_DEV void put(int at, uint2 v) { base[(size_t)at * GK_WPOOL] = v; }
};

template <int FMT>
__global__ void __launch_bounds__(128, GK_WPOOL_MINB) k_trace_wpool(TraceScene ts, PathView V, int qi, size_t half, uint2* __restrict__ stack_mem, uint32_t stop_below) {
    static_assert(FMT == 4, "the pool tracers walk Node4 trees");
    __shared__ WPoolSmem smem[4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WPoolSmem& sm = smem[warp];
    const uint32_t n = V.counts[qi];
    if (wave_stopped(V, qi, n, stop_below)) return;
    if (blockIdx.x == 0 && threadIdx.x == 0) V.counts[qi ^ 1] = 0;
    const float4* ray_o = V.ray_o + (size_t)qi * half;
    const float4* ray_d = V.ray_d + (size_t)qi * half;
    const uint32_t* slot_pixel = V.slot_pixel[qi];
    uint2* const my_stack = stack_mem + ((size_t)blockIdx.x * 4 + warp) * GK_TRAV_STACK * GK_WPOOL;
    for (int i = lane; i < GK_WPOOL; i += 32) sm.ring[PK_FREE][i] = (uint8_t)i;
    uint32_t head[PK_COUNT], tail[PK_COUNT];  // warp-uniform
#pragma unroll
    for (int k = 0; k < PK_COUNT; ++k) { head[k] = 0; tail[k] = 0; }
    tail[PK_FREE] = GK_WPOOL;
    bool drained = false;
    __syncwarp();
    for (;;) {
        // ---- pick the step kind: the first (in priority order) with a full chunk waiting, else the largest rest
        int kind = -1;
        uint32_t cnt = 0;
        {
            uint32_t best = 0;
#pragma unroll
            for (int k = 0; k < PK_COUNT; ++k) {
                uint32_t a = tail[k] - head[k];
                if (k == PK_FREE && drained) a = 0;
                if (best < 32u && a > best) { best = a; kind = k; }  // the first kind that reaches 32 wins, otherwise the largest
            }
            cnt = min(best, 32u);
        }
        if (kind < 0) break;  // every slot is free and the queue is empty
        uint32_t begin = 0;
#pragma unroll
        for (int k = 0; k < PK_COUNT; ++k)
            if (k == kind) { begin = head[k]; head[k] += cnt; }
        const bool act = (uint32_t)lane < cnt;
        const uint32_t slot = act ? sm.ring[kind][(begin + lane) & 127u] : 0u;
        int next = -1;
        if (kind == PK_FREE) {
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(V.counts + 2, cnt);
            base = __shfl_sync(0xffffffffu, base, 0);
            if (base + cnt >= n) drained = true;
            if (act) {
                const uint32_t q = base + lane;
                next = PK_FREE;
                if (q < n) {
                    const float4 o = __ldcs(ray_o + q), d = __ldcs(ray_d + q);
                    const bool any = (__ldcs(slot_pixel + q) & 0x80000000u) != 0;
                    Trav T;
                    WorldQueue world{ray_o + q, ray_d + q};
                    trav_init<FMT>(T, world, ts, f3(o.x, o.y, o.z), f3(d.x, d.y, d.z), GK_EPS, GK_PT_MAX_TRACE_DISTANCE, any);
                    __stcs(V.hit_tuv + q, f4(T.best.t, 0.0f, 0.0f, __uint_as_float(GK_NO_HIT)));  // the miss record; a hit overwrites it
                    __stcs(V.hit_prim + q, 0u);
                    if (!T.done) {
                        sm.ro[0][slot] = T.ro.x; sm.ro[1][slot] = T.ro.y; sm.ro[2][slot] = T.ro.z;
                        sm.idir[0][slot] = T.idir.x; sm.idir[1][slot] = T.idir.y; sm.idir[2][slot] = T.idir.z;
                        sm.best_t[slot] = T.best.t;
                        sm.cur[slot] = 0u;
                        sm.nbase[slot] = 0xFFFFFFFFu;
                        sm.cinst[slot] = GK_NO_HIT;
                        sm.pack[slot] = any ? PP_ANY : 0u;
                        sm.qslot[slot] = q;
                        next = PK_NODE;
                    }
                }
            }
        } else if (kind == PK_NODE) {
            if (act) {
                Trav T;
                const uint32_t pk = sm.pack[slot], nb = sm.nbase[slot];
                set_space<FMT>(T, f3(sm.ro[0][slot], sm.ro[1][slot], sm.ro[2][slot]), f3(sm.idir[0][slot], sm.idir[1][slot], sm.idir[2][slot]));
                T.tmin = GK_EPS;
                T.best.t = sm.best_t[slot];
                T.cur = sm.cur[slot];
                T.cur_inst = 0;  // only feeds the statistics build
                T.sp = (int)(pk & 0xFFu);
                T.nodes = reinterpret_cast<const uint4*>(nb == 0xFFFFFFFFu ? ts.tlas : ts.blas + nb);
                WPoolStack stack{my_stack + slot};
                if (trav_node_core4(T, stack)) {
                    sm.cur[slot] = T.cur;
                    next = !(T.cur & GK_LEAF_BIT) ? PK_NODE : (nb == 0xFFFFFFFFu ? PK_INST : PK_TRI);
                } else {
                    next = PK_POP;
                }
                sm.pack[slot] = (pk & ~0xFFu) | (uint32_t)T.sp;
            }
        } else {
            if (act) {
                Trav T;
                uint32_t pk = sm.pack[slot];
                const uint32_t q = sm.qslot[slot];
                WorldQueue world{ray_o + q, ray_d + q};
                T.cur = sm.cur[slot];
                T.cur_inst = sm.cinst[slot];
                T.sp = (int)(pk & 0xFFu);
                T.sp_inst = (int)((pk >> 8) & 0xFFu);
                T.best.t = sm.best_t[slot];
                T.done = false;
                bool pop = true, space_changed = false;
                if (kind == PK_INST) {
                    if (trav_enter_instance<FMT>(T, world, ts, T.cur & 0x0FFFFFFFu)) {
                        sm.nbase[slot] = __ldg(&ts.inst[T.cur_inst].node_base);
                        sm.sh[0][slot] = T.sh.Sx; sm.sh[1][slot] = T.sh.Sy; sm.sh[2][slot] = T.sh.Sz;
                        sm.shk[slot] = (uint32_t)T.sh.kx | ((uint32_t)T.sh.ky << 2) | ((uint32_t)T.sh.kz << 4);
                        sm.cur[slot] = 0u;
                        pop = false;
                        space_changed = true;
                        next = PK_NODE;
                    }
                } else if (kind == PK_TRI) {
                    const float3 ro = f3(sm.ro[0][slot], sm.ro[1][slot], sm.ro[2][slot]);
                    RayShear sh;
                    const uint32_t k3 = sm.shk[slot];
                    sh.kx = (int)(k3 & 3u); sh.ky = (int)((k3 >> 2) & 3u); sh.kz = (int)((k3 >> 4) & 3u);
                    sh.Sx = sm.sh[0][slot]; sh.Sy = sm.sh[1][slot]; sh.Sz = sm.sh[2][slot];
                    const float4* tris = ts.tris + (size_t)__ldg(&ts.inst[T.cur_inst].tri_base) * 3;
                    const uint32_t first = T.cur & 0x0FFFFFFFu, count = ((T.cur >> 28) & 7u) + 1u;
                    for (uint32_t i = 0; i < count; ++i) {
                        const float4* tp = tris + (size_t)(first + i) * 3;
                        const float4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2);
                        float t, b1, b2;
                        if (!tri_test(f3(a.x, a.y, a.z), f3(b.x, b.y, b.z), f3(c.x, c.y, c.z), ro, sh, t, b1, b2)) continue;
                        if (!(t > GK_EPS && t < GK_PT_MAX_TRACE_DISTANCE)) continue;
                        const uint32_t p = __float_as_uint(a.w);
                        if (pk & PP_ANY) {  // first hit ends the query
                            __stcs(V.hit_tuv + q, f4(t, 0.0f, 0.0f, __uint_as_float(T.cur_inst)));
                            __stcs(V.hit_prim + q, p);
                            T.done = true;
                            break;
                        }
                        bool better = !(pk & PP_HAS_HIT) || t < T.best.t;
                        if (!better && t == T.best.t) {  // tie: (instance, primitive) order; the record of the best hit is in global memory
                            const uint32_t bi = __float_as_uint(__ldcg(&V.hit_tuv[q].w)), bp = __ldcg(V.hit_prim + q);
                            better = T.cur_inst < bi || (T.cur_inst == bi && p < bp);
                        }
                        if (better) {
                            T.best.t = t;
                            pk |= PP_HAS_HIT;
                            __stcg(V.hit_tuv + q, f4(t, b1, b2, __uint_as_float(T.cur_inst)));
                            __stcg(V.hit_prim + q, p);
                        }
                    }
                    sm.best_t[slot] = T.best.t;
                    if (T.done) { pop = false; next = PK_FREE; }
                }
                if (pop) {
                    const uint32_t inst_before = T.cur_inst;
                    WPoolStack stack{my_stack + slot};
                    trav_pop<FMT>(T, world, ts, stack);
                    if (T.done) next = PK_FREE;
                    else next = !(T.cur & GK_LEAF_BIT) ? PK_NODE : (T.cur_inst == GK_NO_HIT ? PK_INST : PK_TRI);
                    space_changed = T.cur_inst != inst_before;  // left the instance: back in world space
                    if (space_changed) sm.nbase[slot] = 0xFFFFFFFFu;
                    sm.cur[slot] = T.cur;
                }
                if (space_changed) {
                    sm.ro[0][slot] = T.ro.x; sm.ro[1][slot] = T.ro.y; sm.ro[2][slot] = T.ro.z;
                    sm.idir[0][slot] = T.idir.x; sm.idir[1][slot] = T.idir.y; sm.idir[2][slot] = T.idir.z;
                    sm.cinst[slot] = T.cur_inst;
                }
                sm.pack[slot] = (pk & ~0xFFFFu) | (uint32_t)T.sp | ((uint32_t)T.sp_inst << 8);
            }
        }
        // ---- append the slots to the queues of their next step: the queues are private to the warp, a ballot per kind is all it takes
#pragma unroll
        for (int k = 0; k < PK_COUNT; ++k) {
            const unsigned m = __ballot_sync(0xffffffffu, next == k);
            if (next == k) sm.ring[k][(tail[k] + __popc(m & ((1u << lane) - 1u))) & 127u] = (uint8_t)slot;
            tail[k] += __popc(m);
        }
        __syncwarp();
    }
}

// The first version: a warp takes 32 rays, every lane runs the simple per-ray loop, the warp refills when all are done.
__global__ void __launch_bounds__(128) k_trace_simple(TraceScene ts, PathView V, int qi, size_t half, uint32_t stop_below) {
    const uint32_t n = V.counts[qi];
    if (wave_stopped(V, qi, n, stop_below)) return;
    if (blockIdx.x == 0 && threadIdx.x == 0) V.counts[qi ^ 1] = 0;
    const int lane = threadIdx.x & 31;
    const float4* ray_o = V.ray_o + (size_t)qi * half;
    const float4* ray_d = V.ray_d + (size_t)qi * half;
    for (;;) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(V.counts + 2, 32u);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= n) return;
        uint32_t slot = base + lane;
        if (slot < n) {
            float4 o = ray_o[slot], d = ray_d[slot];
            Hit h;
#ifdef GK_TRAV_STATS
            {   // diagnostics build (scripts/trav_stats.py): per-frame traversal statistics in counts[32..], see fetch_ray_counters
                LocalStack stack; Trav T; WorldRegs world;
                const bool any = (V.slot_pixel[qi][slot] & 0x80000000u) != 0;
                if (ts.fmt == 8) {
                    trav_init<8>(T, world, ts, f3(o.x, o.y, o.z), f3(d.x, d.y, d.z), GK_EPS, GK_PT_MAX_TRACE_DISTANCE, any);
                    while (!T.done) {
                        if (!(T.cur & GK_LEAF_BIT)) trav_node_step<8>(T, world, ts, stack);
                        else trav_leaf_step<8>(T, world, ts, stack);
                    }
                } else {
                    trav_init<4>(T, world, ts, f3(o.x, o.y, o.z), f3(d.x, d.y, d.z), GK_EPS, GK_PT_MAX_TRACE_DISTANCE, any);
                    while (!T.done) {
                        if (!(T.cur & GK_LEAF_BIT)) trav_node_step<4>(T, world, ts, stack);
                        else trav_leaf_step<4>(T, world, ts, stack);
                    }
                }
                h = T.best;
                unsigned long long* g = reinterpret_cast<unsigned long long*>(V.counts + 32);
                for (int k = 0; k < 8; ++k) atomicAdd(g + k + (any ? 8 : 0), (unsigned long long)T.st[k]);
            }
#else
            if (V.slot_pixel[qi][slot] & 0x80000000u) trace_ray<true>(ts, f3(o.x, o.y, o.z), f3(d.x, d.y, d.z), GK_EPS, GK_PT_MAX_TRACE_DISTANCE, h);
            else trace_ray<false>(ts, f3(o.x, o.y, o.z), f3(d.x, d.y, d.z), GK_EPS, GK_PT_MAX_TRACE_DISTANCE, h);
#endif
            V.hit_tuv[slot] = f4(h.t, h.u, h.v, __uint_as_float(h.inst));
            V.hit_prim[slot] = h.prim;
        }
    }
}

#ifndef GK_SHADE_PREFETCH
#define GK_SHADE_PREFETCH 0
#endif
#ifndef GK_SHADE_GROUP
#define GK_SHADE_GROUP 0
#endif
template <int LEVEL>
GK_DEV void gk_prefetch(const void* p) {
    if (LEVEL == 2) asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
    else asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
// Shade: one thread per queue slot: consume the hit, advance the path to its next ray, compact the survivors into the other queue.
template <int MINB>
__global__ void __launch_bounds__(128, MINB) k_shade(SceneView S_, const __grid_constant__ GkUniformBufferObject cam, RTView R, PathView V, FrameCfg cfg, int qi, unsigned long long* ray_counters) {
    SceneView S = S_;
    S.cam = &cam;
    // The slot's records are requested BEFORE the queue size: every block used to start by waiting a full memory round trip for the 16-byte
    // header (13 % of the kernel's stall samples, profiles/r02_experiments.md) and only then issue its loads. The grid never exceeds the
    // queue's capacity, so a slot beyond the current size is allocated memory with stale contents that are simply not used.
#if GK_SHADE_GROUP
    // The 128 slots of the block are dealt to its threads misses first, hits after (a stable partition on the hit record's instance word):
    // a warp then runs either the sky lookup or the hit resolve + material fetch instead of both (two thirds of the bench frame's slots are
    // misses, so nearly every warp of the slot-ordered kernel executes both paths). The block's records are the same cache lines whichever
    // thread reads them; they are requested (prefetch into L1) before the partition, whose key needs the hit record's round trip.
    const uint32_t slot0 = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t slot = slot0;
    {
        __shared__ uint16_t s_perm[128];
        __shared__ uint32_t s_cnt[2][4];
        const bool cap0 = slot0 < (uint32_t)V.pixels;
        uint32_t inst0 = GK_NO_HIT;
        if (cap0) {
            const size_t i0 = (size_t)qi * V.pixels + slot0;
            gk_prefetch<2>(V.slot_pixel[qi] + slot0); gk_prefetch<2>(V.hit_prim + slot0);
            gk_prefetch<2>(V.ray_o + i0); gk_prefetch<2>(V.ray_d + i0); gk_prefetch<2>(V.rng + i0); gk_prefetch<2>(V.ray_color + i0);
            inst0 = __float_as_uint(V.hit_tuv[slot0].w);
        }
        const uint32_t n0 = V.counts[qi];
        const bool valid0 = cap0 && slot0 < n0;
        const bool miss0 = valid0 && inst0 == GK_NO_HIT, hit0 = valid0 && inst0 != GK_NO_HIT;
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        const unsigned bm = __ballot_sync(0xffffffffu, miss0), bh = __ballot_sync(0xffffffffu, hit0);
        if (lane == 0) { s_cnt[0][w] = __popc(bm); s_cnt[1][w] = __popc(bh); }
        __syncthreads();
        uint32_t miss_before = 0, hit_before = 0, miss_total = 0, hit_total = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (k < w) { miss_before += s_cnt[0][k]; hit_before += s_cnt[1][k]; }
            miss_total += s_cnt[0][k]; hit_total += s_cnt[1][k];
        }
        const unsigned below = (1u << lane) - 1u;
        uint32_t dst;
        if (miss0) dst = miss_before + __popc(bm & below);
        else if (hit0) dst = miss_total + hit_before + __popc(bh & below);
        else {  // slots beyond the queue keep the rest, in order
            const unsigned bi = ~(bm | bh);
            uint32_t inv_before = (uint32_t)w * 32u - miss_before - hit_before;
            dst = miss_total + hit_total + inv_before + __popc(bi & below);
        }
        s_perm[dst] = (uint16_t)threadIdx.x;
        __syncthreads();
        slot = blockIdx.x * blockDim.x + s_perm[threadIdx.x];
    }
#else
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
#endif
    const bool in_cap = slot < (uint32_t)V.pixels;
    uint32_t sp = 0, hprim = 0;
    float4 hv = f4(0, 0, 0, 0);
    SlotRaw raw{};
    if (in_cap) {
#if GK_SHADE_LDCS
        sp = __ldcs(V.slot_pixel[qi] + slot);
        hv = __ldcs(V.hit_tuv + slot);
        hprim = __ldcs(V.hit_prim + slot);
#else
        sp = V.slot_pixel[qi][slot];
        hv = V.hit_tuv[slot];
        hprim = V.hit_prim[slot];
#endif
        raw = load_slot_raw(V, qi, slot);
    }
    const uint4 hdr = *reinterpret_cast<const uint4*>(V.counts);  // queue sizes, trace cursor, stop flag: one load
    if (hdr.w) return;  // the wavefront has stopped (wave_stopped): the queue is k_tail's
    const uint32_t n = qi ? hdr.y : hdr.x;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        V.counts[2] = 0;  // trace cursor of the next wave (this wave's k_trace is done)
        if (n) { V.counts[12] += n; V.counts[13] += 1; }  // rays and launches of the wavefront part, for the statistics
    }
    PixelOut O{R, V.radiance, V.direct, (size_t)R.w * R.h};
    uint32_t nclosest = 0, nany = 0;
    bool want = false;
    int req = REQ_DONE;
    uint32_t pixel = 0;
    PathRegs P;
    if (slot < n && in_cap) {
#if GK_SHADE_PREFETCH
        {   // The per-pixel records the end of a sample touches (two of three slots end one): requested now, they arrive while the hit is
            // being resolved instead of costing a dependent HBM round trip each at the end of the chain. 1: into L2, 2: into L1.
            const uint32_t px = sp & 0x7FFFFFFFu, fl = __float_as_uint(raw.o.w);
            const float4* acc = (fl & FL_HIT_REFLECT) ? V.final_refl + px : V.final_color + px;
            gk_prefetch<GK_SHADE_PREFETCH>(acc);
            gk_prefetch<GK_SHADE_PREFETCH>(V.prim_nrm_cnt + px);
            gk_prefetch<GK_SHADE_PREFETCH>(V.prim_pos_mat + px);
            if (fl & FL_HIT_METAL) gk_prefetch<GK_SHADE_PREFETCH>(V.prim_albedo + px);
        }
#endif
        Hit h;
        h.t = hv.x; h.u = hv.y; h.v = hv.z; h.inst = __float_as_uint(hv.w); h.prim = hprim;
        load_slot(V, S, qi, slot, (sp & 0x80000000u) != 0, P, &raw);
        pixel = sp & 0x7FFFFFFFu;
        if (sp & 0x80000000u) nany++; else nclosest++;
        req = advance(S, O, cfg, V, pixel, P, h);
        want = req != REQ_DONE;
    }
    enqueue(V, qi ^ 1, want, pixel, req, P);
    for (int o = 16; o > 0; o >>= 1) {
        nclosest += __shfl_xor_sync(0xffffffffu, nclosest, o);
        nany += __shfl_xor_sync(0xffffffffu, nany, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (nclosest) atomicAdd(ray_counters + 1, (unsigned long long)nclosest);
        if (nany) atomicAdd(ray_counters + 2, (unsigned long long)nany);
    }
}

// GKCU_TAIL_SORT: the tail's slots ordered by how many samples their pixel still has to run, longest first, so that the paths of a warp
// end at about the same time and the longest ones start in the first blocks (counting sort: 128 buckets, three small kernels).
GK_DEV uint32_t tail_key(const PathView& V, int qi, uint32_t slot) {
    const uint32_t pixel = V.slot_pixel[qi][slot] & 0x7FFFFFFFu;
    const uint32_t sample = __float_as_uint(V.ray_d[(size_t)qi * V.pixels + slot].w) & 0xFFFFu;
    const uint32_t end = __float_as_uint(V.prim_nrm_cnt[pixel].w) & 0xFFFFu;
    const uint32_t rem = end > sample ? end - sample : 0u;
    return 127u - min(rem, 127u);
}
__global__ void k_tail_hist(PathView V, int qi, uint32_t* __restrict__ hist) {
    if (V.counts[3]) qi = (int)V.counts[3] - 1;
    const uint32_t n = V.counts[qi];
    __shared__ uint32_t s_h[128];
    if (threadIdx.x < 128) s_h[threadIdx.x] = 0;
    __syncthreads();
    for (uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x; slot < n; slot += gridDim.x * blockDim.x) atomicAdd(&s_h[tail_key(V, qi, slot)], 1u);
    __syncthreads();
    if (threadIdx.x < 128 && s_h[threadIdx.x]) atomicAdd(&hist[threadIdx.x], s_h[threadIdx.x]);
}
__global__ void k_tail_scan(uint32_t* __restrict__ hist) {  // hist[0..127] counts -> hist[128..255] first position of each bucket
    if (threadIdx.x == 0) {
        uint32_t acc = 0;
        for (int k = 0; k < 128; ++k) { hist[128 + k] = acc; acc += hist[k]; }
    }
}
__global__ void k_tail_scatter(PathView V, int qi, uint32_t* __restrict__ hist, uint32_t* __restrict__ order) {
    if (V.counts[3]) qi = (int)V.counts[3] - 1;
    const uint32_t n = V.counts[qi];
    for (uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x; slot < n; slot += gridDim.x * blockDim.x)
        order[atomicAdd(&hist[128 + tail_key(V, qi, slot)], 1u)] = slot;
}

#ifndef GK_TAIL_VOTE
#define GK_TAIL_VOTE 0
#endif
// Tail / megakernel: every thread runs its pixel's program to completion with the traversal inlined.
#ifndef GK_TAIL_MINB
#define GK_TAIL_MINB 4  // 4 resident blocks per SM = at most 128 registers (127 used); 1 lets the compiler take 149 and the tail 3.7 ms instead of 2.7; 5 / 6 / 8 (96 / 80 / 64 registers): 2.76 / 3.22 / 3.41 ms
#endif
__global__ void __launch_bounds__(128, GK_TAIL_MINB) k_tail(SceneView S_, const __grid_constant__ GkUniformBufferObject cam, RTView R, PathView V, FrameCfg cfg, int qi, unsigned long long* ray_counters, int after_wavefront, const uint32_t* __restrict__ order, int spread, int solo) {
    SceneView S = S_;
    S.cam = &cam;
    if (after_wavefront && V.counts[3]) qi = (int)V.counts[3] - 1;  // the queue the device stopped the wavefront at (wave_stopped)
    const uint32_t n = V.counts[qi];
    PixelOut O{R, V.radiance, V.direct, (size_t)R.w * R.h};
    const size_t half = (size_t)R.w * R.h;
    uint32_t nclosest = 0, nany = 0;
#if GK_TAIL_VOTE
    // The rays of a warp are traced together (trace_ray_warp: the warp votes on the step kind), then the lanes that had one shade together.
    // The trip counts are warp-uniform; a lane without a slot or with a finished path just sits the calls out.
    for (uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x; __any_sync(0xffffffffu, slot < n); slot += gridDim.x * blockDim.x) {
        const bool have = slot < n;
        uint32_t sp = have ? V.slot_pixel[qi][slot] : 0u;
        uint32_t pixel = sp & 0x7FFFFFFFu;
        PathRegs P;
        if (have) load_slot(V, S, qi, slot, (sp & 0x80000000u) != 0, P);
        else { P.ro = f3(0, 0, 0); P.rd = f3(0, 0, 1); }
        int req = !have ? REQ_DONE : ((sp & 0x80000000u) ? REQ_ANY : REQ_CLOSEST);
        while (__any_sync(0xffffffffu, req != REQ_DONE)) {
            Hit h;
            trace_ray_warp(S.ts, P.ro, P.rd, GK_EPS, GK_PT_MAX_TRACE_DISTANCE, req == REQ_ANY, req != REQ_DONE, h);
            if (req != REQ_DONE) {
                if (req == REQ_ANY) nany++; else nclosest++;
                req = advance(S, O, cfg, V, pixel, P, h);
            }
        }
    }
#else
    // Ordered tail: rank r of the longest-first order goes to warp r mod W, so the W longest paths sit in W different warps (the long paths of
    // a frame are neighbours on screen - the pixels of a glass ball - and in queue order they share warps, where every bounce of one waits
    // for the divergent bounces of the others) and a warp holds about `spread` paths.
    uint32_t at0 = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
    if (order) {
        const uint32_t grid_warps = gridDim.x * (blockDim.x >> 5);
        const uint32_t K = min((uint32_t)solo, min(n, grid_warps / 2u));  // the K longest paths get a warp each
        const uint32_t rest = n - K;
        const uint32_t W = max(1u, min(grid_warps - K, (rest + (uint32_t)spread - 1u) / (uint32_t)spread));
        const uint32_t w = at0 >> 5, lane = threadIdx.x & 31u;
        if (w < K) { at0 = lane == 0 ? w : n; stride = n; }
        else if (w - K >= W) at0 = n;  // this warp has no part in the ordered tail
        else { at0 = K + lane * W + (w - K); stride = 32u * W; }
    }
    for (uint32_t at = at0; at < n; at += stride) {
        const uint32_t slot = order ? order[at] : at;
        uint32_t sp = V.slot_pixel[qi][slot];
        uint32_t pixel = sp & 0x7FFFFFFFu;
        PathRegs P;
        load_slot(V, S, qi, slot, (sp & 0x80000000u) != 0, P);
        int req = (sp & 0x80000000u) ? REQ_ANY : REQ_CLOSEST;
        while (req != REQ_DONE) {
            Hit h;
            if (req == REQ_ANY) { trace_ray<true>(S.ts, P.ro, P.rd, GK_EPS, GK_PT_MAX_TRACE_DISTANCE, h); nany++; }
            else { trace_ray<false>(S.ts, P.ro, P.rd, GK_EPS, GK_PT_MAX_TRACE_DISTANCE, h); nclosest++; }
            req = advance(S, O, cfg, V, pixel, P, h);
        }
    }
#endif
    for (int o = 16; o > 0; o >>= 1) {
        nclosest += __shfl_xor_sync(0xffffffffu, nclosest, o);
        nany += __shfl_xor_sync(0xffffffffu, nany, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (nclosest) atomicAdd(ray_counters + 1, (unsigned long long)nclosest);
        if (nany) atomicAdd(ray_counters + 2, (unsigned long long)nany);
    }
}

// The tail in segments: every thread advances its path by at most `max_rays` rays (0 = to the end), and the paths that are still alive are
// compacted into the other queue for the next segment, so the warps of a segment start full again instead of thinning out to the few
// longest paths (k_tail runs at 3.9 of 32 lanes). Segment `seg` reads queue (q0 + seg) & 1, q0 being the queue the wavefront stopped at, and
// counts the entries it leaves behind in counts[17 + seg] (all zero at the start of the frame).
__global__ void __launch_bounds__(128) k_tail_seg(SceneView S_, const __grid_constant__ GkUniformBufferObject cam, RTView R, PathView V, FrameCfg cfg, int qi, unsigned long long* ray_counters,
                                                  int seg, int max_rays) {
    SceneView S = S_;
    S.cam = &cam;
    const int q0 = V.counts[3] ? (int)V.counts[3] - 1 : qi;
    const int qin = (q0 + seg) & 1, qout = qin ^ 1;
    const uint32_t n = seg == 0 ? V.counts[q0] : V.counts[16 + seg];
    uint32_t* out_count = V.counts + 17 + seg;
    PixelOut O{R, V.radiance, V.direct, (size_t)R.w * R.h};
    uint32_t nclosest = 0, nany = 0;
    const int lane = threadIdx.x & 31;
    for (uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x; __any_sync(0xffffffffu, slot < n); slot += gridDim.x * blockDim.x) {
        const bool have = slot < n;
        uint32_t pixel = 0;
        int req = REQ_DONE;
        PathRegs P;
        if (have) {
            const uint32_t sp = V.slot_pixel[qin][slot];
            pixel = sp & 0x7FFFFFFFu;
            load_slot(V, S, qin, slot, (sp & 0x80000000u) != 0, P);
            req = (sp & 0x80000000u) ? REQ_ANY : REQ_CLOSEST;
            for (int rays = 0; req != REQ_DONE && (max_rays == 0 || rays < max_rays); ++rays) {
                Hit h;
                if (req == REQ_ANY) { trace_ray<true>(S.ts, P.ro, P.rd, GK_EPS, GK_PT_MAX_TRACE_DISTANCE, h); nany++; }
                else { trace_ray<false>(S.ts, P.ro, P.rd, GK_EPS, GK_PT_MAX_TRACE_DISTANCE, h); nclosest++; }
                req = advance(S, O, cfg, V, pixel, P, h);
            }
        }
        const bool want = req != REQ_DONE;
        const unsigned mask = __ballot_sync(0xffffffffu, want);
        if (mask) {
            const int leader = __ffs(mask) - 1;
            uint32_t base = 0;
            if (lane == leader) base = atomicAdd(out_count, (uint32_t)__popc(mask));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (want) store_slot(V, qout, base + __popc(mask & ((1u << lane) - 1u)), pixel, req, P);
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        nclosest += __shfl_xor_sync(0xffffffffu, nclosest, o);
        nany += __shfl_xor_sync(0xffffffffu, nany, o);
    }
    if (lane == 0) {
        if (nclosest) atomicAdd(ray_counters + 1, (unsigned long long)nclosest);
        if (nany) atomicAdd(ray_counters + 2, (unsigned long long)nany);
    }
}

// The tail as persistent warps that schedule by vote: every lane owns one path and is either walking its pending ray (node step /
// instance entry / triangle leaf, as in k_trace) or waiting to shade its hit; each iteration the warp runs the one step kind most of its
// lanes wait for, so a lane does not wait for the longest ray of its warp before it shades (k_tail), and lanes whose path has ended pull the
// next slot of the tail queue. Same per-path sequence of rays and the same `advance`, so the same bits. counts[15] is the queue cursor
// (zero at the start of the frame).
#ifndef GK_FUSED_REFILL
#define GK_FUSED_REFILL 8  // idle lanes that make the warp go back for more paths while the queue has some
#endif
__global__ void __launch_bounds__(128, 4) k_tail_fused(SceneView S_, const __grid_constant__ GkUniformBufferObject cam, RTView R, PathView V, FrameCfg cfg, int qi,
                                                       unsigned long long* ray_counters, int after_wavefront) {
    SceneView S = S_;
    S.cam = &cam;
    if (after_wavefront && V.counts[3]) qi = (int)V.counts[3] - 1;
    const uint32_t n = V.counts[qi];
    PixelOut O{R, V.radiance, V.direct, (size_t)R.w * R.h};
    const int lane = threadIdx.x & 31;
    uint32_t nclosest = 0, nany = 0;
    LocalStack stack;
    Trav T;
    WorldRegs world;
    PathRegs P;
    uint32_t pixel = 0;
    int req = REQ_DONE;
    bool has_path = false, drained = false;
    uint32_t guard = 0;
    T.done = true; T.cur = 0; T.cur_inst = GK_NO_HIT;
    for (;;) {
        if (!drained) {
            const unsigned idle = __ballot_sync(0xffffffffu, !has_path);
            if (idle) {
                const int leader = __ffs(idle) - 1;
                uint32_t base = 0;
                if (lane == leader) base = atomicAdd(V.counts + 15, (uint32_t)__popc(idle));
                base = __shfl_sync(0xffffffffu, base, leader);
                if (!has_path) {
                    const uint32_t slot = base + __popc(idle & ((1u << lane) - 1u));
                    if (slot < n) {
                        const uint32_t sp = V.slot_pixel[qi][slot];
                        pixel = sp & 0x7FFFFFFFu;
                        load_slot(V, S, qi, slot, (sp & 0x80000000u) != 0, P);
                        req = (sp & 0x80000000u) ? REQ_ANY : REQ_CLOSEST;
                        trav_init<4>(T, world, S.ts, P.ro, P.rd, GK_EPS, GK_PT_MAX_TRACE_DISTANCE, req == REQ_ANY);
                        has_path = true;
                    }
                }
                if (base + (uint32_t)__popc(idle) >= n) drained = true;
            }
        }
        if (!__any_sync(0xffffffffu, has_path)) break;
        for (;;) {
            const bool tracing = has_path && !T.done;
            const bool leaf = (T.cur & GK_LEAF_BIT) != 0;
            const bool wN = tracing && !leaf;
            const bool wI = tracing && leaf && T.cur_inst == GK_NO_HIT;
            const bool wT = tracing && leaf && T.cur_inst != GK_NO_HIT;
            const bool wS = has_path && T.done;
            const int nN = __popc(__ballot_sync(0xffffffffu, wN)), nI = __popc(__ballot_sync(0xffffffffu, wI)), nT = __popc(__ballot_sync(0xffffffffu, wT)),
                      nS = __popc(__ballot_sync(0xffffffffu, wS));
            const int busy = nN + nI + nT + nS;
            if (busy == 0) break;
            if (++guard > (1u << 27)) { if (lane == 0) V.counts[14] = 0xDEADu; return; }  // cannot happen; a scheduling bug must not hang the device
            if (!drained && 32 - busy >= GK_FUSED_REFILL) break;  // enough idle lanes: fetch more paths
            if (nN >= nT && nN >= nI && nN >= nS) {
                if (wN) trav_node_step<4>(T, world, S.ts, stack);
            } else if (nT >= nI && nT >= nS) {
                if (wT) trav_leaf_step<4>(T, world, S.ts, stack);
            } else if (nI >= nS) {
                if (wI) trav_leaf_step<4>(T, world, S.ts, stack);
            } else if (wS) {
                const Hit h = T.best;
                if (req == REQ_ANY) nany++; else nclosest++;
                req = advance(S, O, cfg, V, pixel, P, h);
                if (req == REQ_DONE) has_path = false;
                else trav_init<4>(T, world, S.ts, P.ro, P.rd, GK_EPS, GK_PT_MAX_TRACE_DISTANCE, req == REQ_ANY);
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        nclosest += __shfl_xor_sync(0xffffffffu, nclosest, o);
        nany += __shfl_xor_sync(0xffffffffu, nany, o);
    }
    if (lane == 0) {
        if (nclosest) atomicAdd(ray_counters + 1, (unsigned long long)nclosest);
        if (nany) atomicAdd(ray_counters + 2, (unsigned long long)nany);
    }
}

// Multi-GPU resolve: SINGLE_* = sums / S + direct (Shading.slang:1175-1183) after the NCCL reduce.
__global__ void k_resolve(RTView R, const float4* __restrict__ radiance, const float4* __restrict__ direct, const uint16_t* __restrict__ moment,
                          const GkUniformBufferObject* __restrict__ C) {
    uint32_t pixel = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t n = R.w * R.h;
    if (pixel >= n) return;
    int sampleMultiplier = moment[pixel] > 0 ? 4 : 1;
    int sample = max(1, (int)C->NumberOfSamples * sampleMultiplier);
    float4 fc = radiance[pixel] / float(sample), fr = radiance[(size_t)n + pixel] / float(sample);
    fc = fc + direct[pixel];
    store_h4(R.single_diffuse, pixel, fc);
    store_h4(R.single_specular, pixel, fr);
}

// Fused multi-GPU reduce + resolve of one strip of rows: peer loads of every rank's un-normalised sums, rank-ordered sum,
// Shading.slang:1175-1183, peer stores of the fp16 results into rank 0's SINGLE_* images. 32 B in per rank and pixel (+ 16 B direct term), 16 B out.
struct PeerRadiance { const float4* rad[GKCU_MAX_PEERS]; };
__global__ void __launch_bounds__(256) k_resolve_strip(uint2* __restrict__ out_diffuse, uint2* __restrict__ out_specular, PeerRadiance P, int world,
                                                       const float4* __restrict__ direct0, const uint16_t* __restrict__ moment,
                                                       const __grid_constant__ GkUniformBufferObject cam, uint32_t n, uint32_t first, uint32_t last) {
    const uint32_t pixel = first + blockIdx.x * blockDim.x + threadIdx.x;
    if (pixel >= last || pixel >= n) return;
    float4 fc = __ldcs(P.rad[0] + pixel), fr = __ldcs(P.rad[0] + (size_t)n + pixel);
    for (int r = 1; r < world; ++r) {
        fc = fc + __ldcs(P.rad[r] + pixel);
        fr = fr + __ldcs(P.rad[r] + (size_t)n + pixel);
    }
    const int sampleMultiplier = moment[pixel] > 0 ? 4 : 1;
    const int sample = max(1, (int)cam.NumberOfSamples * sampleMultiplier);
    fc = fc / float(sample);
    fr = fr / float(sample);
    fc = fc + __ldcs(direct0 + pixel);
    store_h4(out_diffuse, pixel, fc);
    store_h4(out_specular, pixel, fr);
}

// Device-side cross-GPU barrier for the strip resolve. Each rank owns a flag block {ready[8], done[8], timeouts}; rank r stores its
// frame counter into slot r of EVERY rank's block (peer stores, release at system scope: by cumulativity everything this
// stream wrote before - the radiance sums, or the resolved strip - is visible to whoever acquires the flag), and waits until
// all slots of its own block have reached the counter.
struct PeerFlags { unsigned long long* block[GKCU_MAX_PEERS]; };
__global__ void k_peer_signal(PeerFlags F, int world, int rank, int slot0, unsigned long long epoch) {
    const int r = threadIdx.x;
    if (r >= world) return;
    __threadfence_system();
    unsigned long long* p = F.block[r] + slot0 + rank;
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(epoch) : "memory");
}
__global__ void k_peer_wait(unsigned long long* mine, int world, int slot0, unsigned long long epoch) {
    const int r = threadIdx.x;
    if (r >= world) return;
    const unsigned long long* p = mine + slot0 + r;
    const long long t0 = clock64();
    for (;;) {
        unsigned long long v;
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
        if (v >= epoch) break;
        if (clock64() - t0 > 4000000000ll) {  // ~2 s: a peer is gone; do not hang the GPU
            atomicAdd(mine + 16, 1ull);
            break;
        }
        __nanosleep(100);
    }
}

__global__ void k_trace_rays(TraceScene ts, const GkcuRay* __restrict__ rays, uint32_t n, GkcuHit* __restrict__ hits, int any_hit) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    GkcuRay r = rays[i];
    Hit h;
    if (any_hit) trace_ray<true>(ts, f3(r.ox, r.oy, r.oz), f3(r.dx, r.dy, r.dz), r.tmin, r.tmax, h);
    else trace_ray<false>(ts, f3(r.ox, r.oy, r.oz), f3(r.dx, r.dy, r.dz), r.tmin, r.tmax, h);
    GkcuHit o;
    o.t = any_hit ? 0.0f : h.t; o.u = h.u; o.v = h.v; o.inst = any_hit ? (h.inst != GK_NO_HIT ? 0u : GK_NO_HIT) : h.inst; o.prim = any_hit ? 0u : h.prim;
    if (!any_hit && h.inst == GK_NO_HIT) { o.t = 0; o.u = 0; o.v = 0; o.prim = 0; }
    hits[i] = o;
}

}  // namespace

SceneView make_scene_view(const Context& c) {
    SceneView S;
    S.ts = make_trace_scene(c);
    S.vertices = c.vertices.as<uint2>();
    S.indices = c.indices.as<uint32_t>();
    S.offsets = c.offsets.as<GkModelData>();
    S.materials = c.materials.as<GkMaterial>();
    S.material_count = c.material_count;
    S.nodes = c.nodes.as<GkNodeProxy>();
    S.shade_tris = c.shade_tris.as<uint4>();
    S.shade_inst = c.shade_inst.as<uint4>();
    S.textures = c.tex_table.as<TexDesc>();
    S.tex_count = c.tex_count;
    S.probes = c.has_probes ? c.probes.as<ProbeRec>() : nullptr;
    S.probe_meta = c.has_probes ? c.probe_meta.as<ProbeMeta>() : nullptr;
    S.cam = c.ubo.as<GkUniformBufferObject>();
    S.lights = c.light_count ? c.lights.as<GkLightObject>() : nullptr;
    S.sky_sh = c.sky_sh.as<GkSphericalHarmonics>();
    S.sky_sh_count = c.sky_sh_count;
    return S;
}
namespace {

RTView make_rt_view(const Context& c) {
    RTView R;
    const RenderTargets& t = c.rt;
    R.w = t.w; R.h = t.h;
    R.single_diffuse = t.single_diffuse; R.single_specular = t.single_specular; R.albedo = t.albedo; R.normal = t.normal;
    R.objectid0 = t.objectid[t.cur]; R.primitiveid = t.primitiveid; R.motion = t.motion; R.prev_depth = t.prev_depth; R.moment = t.moment;
    R.vis = nullptr;
    return R;
}
// Lane `lane` of the frame: its own pair of queues (capacity lane_cap each) and counters; the per-pixel state is shared.
PathView make_path_view(const Context& c, int lane = 0) {
    PathBuffers& b = *c.path;
    PathView V;
    const size_t cap = b.lane_cap, q = (size_t)lane * 2 * cap, h = (size_t)lane * cap;
    V.rng = b.rng.as<uint4>() + q;
    V.pixels = cap;
    V.pos_mat = b.pos_mat.as<float4>() + q; V.nrm = b.nrm_flags.as<float4>() + q; V.dir = b.dir_cnt.as<float4>() + q;
    V.ray_color = b.ray_color.as<float4>() + q; V.final_color = b.final_color.as<float4>(); V.final_refl = b.final_refl.as<float4>();
    V.prim_pos_mat = b.prim_pos_mat.as<float4>(); V.prim_nrm_cnt = b.prim_nrm_cnt.as<float4>(); V.prim_albedo = b.prim_albedo.as<float4>();
    V.ray_o = b.ray_o.as<float4>() + q; V.ray_d = b.ray_d.as<float4>() + q;
    V.slot_pixel[0] = b.slot_pixel[0].as<uint32_t>() + h; V.slot_pixel[1] = b.slot_pixel[1].as<uint32_t>() + h;
    V.hit_tuv = b.hit_tuv.as<float4>() + h; V.hit_prim = b.hit_prim.as<uint32_t>() + h;
    V.counts = b.counts.as<uint32_t>() + (size_t)lane * 64;
    V.radiance = b.radiance.as<float4>(); V.direct = b.direct.as<float4>();
    V.order = nullptr;
    V.publish = nullptr;
    return V;
}

}  // namespace

int path_resize(Context& c) {
    if (!c.path) c.path = new PathBuffers();
    PathBuffers& b = *c.path;
    size_t n = (size_t)c.rt.w * c.rt.h;
    b.pixels = n;
    // k_primary deals its 8 x 16 pixel tiles to the lanes as a checkerboard: a lane gets at most one tile more per tile row than its share
    b.lanes = std::max(1, std::min(4, c.lanes));
    const size_t tiles_x = (c.rt.w + 7) / 8, tiles_y = (c.rt.h + 15) / 16;
    b.lane_cap = b.lanes == 1 ? n : ((tiles_x * tiles_y + b.lanes - 1) / b.lanes + tiles_y + 8) * 128;
    const size_t m = b.lane_cap * b.lanes;  // slots over all lanes (>= n)
    bool ok = b.rng.reserve(2 * m * 16) && b.pos_mat.reserve(2 * m * 16) && b.nrm_flags.reserve(2 * m * 16) && b.dir_cnt.reserve(2 * m * 16) &&
              b.ray_color.reserve(2 * m * 16) && b.final_color.reserve(n * 16) && b.final_refl.reserve(n * 16) && b.prim_pos_mat.reserve(n * 16) &&
              b.prim_nrm_cnt.reserve(n * 16) && b.prim_albedo.reserve(n * 16) && b.ray_o.reserve(2 * m * 16) && b.ray_d.reserve(2 * m * 16) &&
              b.slot_pixel[0].reserve(m * 4) && b.slot_pixel[1].reserve(m * 4) && b.hit_tuv.reserve(m * 16) && b.hit_prim.reserve(m * 4) &&
              b.counts.reserve(4 * 64 * 4) && b.tail_order.reserve(m * 4) && b.tail_hist.reserve(4 * 256 * 4) && b.radiance.reserve(2 * n * 16) && b.direct.reserve(n * 16) &&
              b.pool_stack.reserve((size_t)b.lanes * c.sm_count * std::max<size_t>((size_t)GK_POOL_MINB * GK_POOL, (size_t)GK_WPOOL_MINB * 4 * GK_WPOOL) * GK_TRAV_STACK * sizeof(uint2));
    if (c.sort_rays) {
        size_t tmp = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, tmp, (unsigned long long*)nullptr, (unsigned long long*)nullptr, (uint32_t*)nullptr, (uint32_t*)nullptr, (int)b.lane_cap, 0, 64, c.stream);
        b.sort_tmp_bytes = tmp;
        ok = ok && b.sort_keys[0].reserve(m * 8) && b.sort_keys[1].reserve(m * 8) && b.sort_vals[0].reserve(m * 4) && b.sort_vals[1].reserve(m * 4) && b.sort_tmp.reserve(tmp);
    }
    b.pool_stack_stride = (size_t)c.sm_count * std::max<size_t>((size_t)GK_POOL_MINB * GK_POOL, (size_t)GK_WPOOL_MINB * 4 * GK_WPOOL) * GK_TRAV_STACK;
    if (!ok) return 1;
    // ranks other than 0 never write the direct-term plane; keep it defined for a host that resolves on such a context
    GK_CUDA(cudaMemsetAsync(b.direct.p, 0, n * 16, c.stream));
    // k_shade requests its slot's records before it knows the queue size (see there): slots beyond the size are read and not used, so they
    // only have to be defined (compute-sanitizer initcheck)
    GK_CUDA(cudaMemsetAsync(b.rng.p, 0, 2 * m * 16, c.stream));
    GK_CUDA(cudaMemsetAsync(b.ray_o.p, 0, 2 * m * 16, c.stream));
    GK_CUDA(cudaMemsetAsync(b.ray_d.p, 0, 2 * m * 16, c.stream));
    GK_CUDA(cudaMemsetAsync(b.ray_color.p, 0, 2 * m * 16, c.stream));
    GK_CUDA(cudaMemsetAsync(b.slot_pixel[0].p, 0, m * 4, c.stream));
    GK_CUDA(cudaMemsetAsync(b.slot_pixel[1].p, 0, m * 4, c.stream));
    GK_CUDA(cudaMemsetAsync(b.hit_tuv.p, 0, m * 16, c.stream));
    GK_CUDA(cudaMemsetAsync(b.hit_prim.p, 0, m * 4, c.stream));
    b.host_counts = c.pub_host + 64;
    return 0;
}

void path_release(Context& c) {
    if (!c.path) return;
    PathBuffers& b = *c.path;
    for (DevBuf* d : {&b.rng, &b.pos_mat, &b.nrm_flags, &b.dir_cnt, &b.ray_color, &b.final_color, &b.final_refl, &b.prim_pos_mat, &b.prim_nrm_cnt,
                      &b.prim_albedo, &b.ray_o, &b.ray_d, &b.slot_pixel[0], &b.slot_pixel[1], &b.hit_tuv, &b.hit_prim, &b.counts, &b.tail_order, &b.tail_hist, &b.radiance, &b.direct, &b.pool_stack, &b.sort_keys[0], &b.sort_keys[1], &b.sort_vals[0], &b.sort_vals[1], &b.sort_tmp})
        d->release();
    delete c.path;
    c.path = nullptr;
}

int run_rt_pass(Context& c, const GkUniformBufferObject& ubo, const GkcuFrameParams& p) {
    // k_trace<5> is sensitive to L1 capacity (nodes, triangles and the spilled part of the stacks live there): ask for exactly the
    // shared memory its 8 resident blocks need and leave the rest of the 256 KB array to L1.
    if (c.carveout_applied != c.trace_carveout + 2) {  // per context (= per device): cudaFuncSetAttribute applies to the current device
        int pct = c.trace_carveout;
        for (int fmt = 0; fmt < 2; ++fmt) {
            const void* fn = fmt ? (const void*)k_trace<5, 8> : (const void*)k_trace<5, 4>;
            int p = pct;
            if (p < 0) {
                cudaFuncAttributes fa;
                int smem_sm = 0;
                if (cudaFuncGetAttributes(&fa, fn) == cudaSuccess &&
                    cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, c.device) == cudaSuccess && smem_sm > 0)
                    p = std::min(100, (int)((GK_TRACE_MINB * (fa.sharedSizeBytes + 1024) * 100 + smem_sm - 1) / smem_sm));
            }
            if (p >= 0) cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, p);
            // An SM runs one shared-memory configuration at a time: a kernel with another carve-out cannot start on an SM while blocks
            // of k_trace are resident. The other kernels of the rt pass ask for the same split (GKCU_MATCH_CARVEOUT), so that the blocks of
            // a following (or, with several lanes, concurrent) kernel can move in while the last k_trace blocks drain.
            if (p >= 0 && c.match_carveout && fmt == (c.node_format == 8 ? 1 : 0)) {
                for (const void* g : {(const void*)k_shade<1>, (const void*)k_shade<6>, (const void*)k_shade<7>, (const void*)k_shade<8>, (const void*)k_shade<10>,
                                      (const void*)k_tail, (const void*)k_primary})
                    cudaFuncSetAttribute(g, cudaFuncAttributePreferredSharedMemoryCarveout, p);
            }
        }
        c.carveout_applied = c.trace_carveout + 2;
    }
    cudaStream_t st = c.stream;
    PathBuffers& b = *c.path;
    SceneView S = make_scene_view(c);
    RTView R = make_rt_view(c);
    if (p.primary_visibility) R.vis = c.rt.minigbuffer;
    const int lanes = b.lanes;
    PathLanes L{};
    L.n = (uint32_t)lanes;
    for (int l = 0; l < lanes; ++l) L.v[l] = make_path_view(c, l);
    FrameCfg cfg;
    cfg.world = std::max(1u, p.spp_world);
    cfg.rank = std::min(p.spp_rank, cfg.world - 1);
    cfg.frame_seed = cfg.world == 1 ? ubo.TotalFrames : ubo.TotalFrames * cfg.world + cfg.rank;
    cfg.to_radiance = p.skip_post ? 1u : 0u;
    cfg.exit_after_first = ubo.FastGather ? 1u : 0u;
    GK_CUDA(cudaMemsetAsync(b.counts.p, 0, 4 * 64 * 4, st));
    auto ray_counters_of = [&](int l) { return reinterpret_cast<unsigned long long*>(L.v[l].counts + 4); };
    GkcuKernelStats ks{};
    dim3 grid((R.w + 7) / 8, (R.h + 15) / 16);
    cudaEventRecord(c.kev[0], st);
    k_primary<<<grid, 128, 0, st>>>(S, ubo, R, L, cfg, ray_counters_of(0));
    cudaEventRecord(c.kev[1], st);
    c.launches++;
    const uint32_t pixels = R.w * R.h;
    uint64_t iterations = 0;
    // Lane l runs on its own stream (lane 0 of a single-lane frame stays on the context stream): waves are stream-ordered inside a
    // lane and independent across lanes, so the hardware overlaps one lane's k_shade and the draining end of its k_trace with
    // the other lanes' kernels.
    cudaStream_t ls[4];
    for (int l = 0; l < lanes; ++l) ls[l] = lanes == 1 ? st : c.lane_stream[l];
    if (lanes > 1) {
        cudaEventRecord(c.lane_ev[0][3], st);
        for (int l = 0; l < lanes; ++l) cudaStreamWaitEvent(ls[l], c.lane_ev[0][3], 0);
    }
    int qi[4] = {0, 0, 0, 0};
    uint32_t bound[4] = {0, 0, 0, 0};
    bool active[4] = {false, false, false, false};
    uint32_t tail_threshold = 0xFFFFFFFFu;
    if (p.tracer == 0) {
        // Wavefront loop while the wave is wide; k_tail finishes the thin tail in one launch. The DEVICE decides when the wavefront ends:
        // the first k_trace that finds its queue at or below the threshold raises the stop flag (counts[3] = queue + 1) and every later
        // kernel of the loop returns at once, so the host never has to wait for a queue size before it launches the next wave. It still
        // reads the sizes - asynchronously, through a ring of pinned words that trails the launches by up to kRing - 1 waves - for two
        // things: the grids of the following waves (the queue can only shrink: every path has at most one successor, so an old size is
        // an upper bound) and to stop launching once a size at or below the threshold has come back. The GPU therefore always has
        // several waves queued (no launch gaps), and the few waves launched past the stop cost two empty kernels each.
        // hand-over once a wave is narrower than pixels / tail_div, but with no more paths than one residency round of k_tail holds
        // (tail_cap per SM: at 4K pixels / 32 would be 3.5 rounds of it) and no fewer than a block per SM
        tail_threshold = std::max<uint32_t>(std::min<uint32_t>(pixels / (uint32_t)std::max(1, c.tail_div), (uint32_t)c.sm_count * (uint32_t)c.tail_cap),
                                            (uint32_t)c.sm_count * 128u) / (uint32_t)lanes;
        constexpr int kRing = 8;
        const int ahead = std::max(1, std::min(kRing - 1, c.wave_chunk));
        bool any = false;
        int launched[4] = {0, 0, 0, 0}, seen[4] = {0, 0, 0, 0};  // waves launched / waves whose input queue size the host has read
        for (int l = 0; l < lanes; ++l) {
            bound[l] = (uint32_t)std::min<size_t>(pixels, b.lane_cap);  // k_primary queues at most one path per pixel: no need to wait for it
            active[l] = bound[l] > tail_threshold;
            any = any || active[l];
        }
        auto consume = [&](int l, bool block) {  // read the sizes that have arrived (block: wait for the oldest outstanding one first)
            while (seen[l] < launched[l]) {
                const int k = seen[l] % kRing;
                if (block) { GK_CUDA(cudaEventSynchronize(c.pev[l][k])); block = false; }
                else if (cudaEventQuery(c.pev[l][k]) != cudaSuccess) break;
                if (c.kernel_timing) {
                    float a = 0, bms = 0;
                    cudaEventElapsedTime(&a, c.wev[l][3 * k], c.wev[l][3 * k + 1]);
                    cudaEventElapsedTime(&bms, c.wev[l][3 * k + 1], c.wev[l][3 * k + 2]);
                    ks.trace_ms += a;
                    ks.shade_ms += bms;
                }
                const uint32_t size = b.host_counts[64 * l + 20 + k];  // stored by that wave's k_trace (wave_stopped): its input queue, 0 = stopped
                bound[l] = std::min(bound[l], size);
                if (bound[l] <= tail_threshold) active[l] = false;
                ++seen[l];
            }
            return 0;
        };
        while (any) {
            any = false;
            for (int l = 0; l < lanes; ++l) {
                if (!active[l]) continue;
                const PathView& V = L.v[l];
                const size_t half = V.pixels;
                cudaStream_t s2 = ls[l];
                const int k = launched[l] % kRing;
                {
                    PathView Vt = V;
                    Vt.publish = c.pub_dev + 64 + 64 * l + 20 + k;
                    if (c.sort_rays) {
                        const uint32_t nmax = bound[l];
                        const float3 lo = make_float3(c.tlas.root_lo[0], c.tlas.root_lo[1], c.tlas.root_lo[2]);
                        float3 ext = make_float3(c.sort_extent > 0 ? c.sort_extent : c.tlas.root_hi[0] - lo.x, c.sort_extent > 0 ? c.sort_extent : c.tlas.root_hi[1] - lo.y, c.sort_extent > 0 ? c.sort_extent : c.tlas.root_hi[2] - lo.z);
                        float3 lo2 = lo;
                        if (c.sort_extent > 0) lo2 = make_float3(-0.5f * c.sort_extent, -0.5f * c.sort_extent, -0.5f * c.sort_extent);
                        const float3 inv = make_float3(1.0f / std::max(ext.x, 1e-20f), 1.0f / std::max(ext.y, 1e-20f), 1.0f / std::max(ext.z, 1e-20f));
                        const int bits = std::max(1, std::min(19, c.sort_rays));
                        unsigned long long *k0 = b.sort_keys[0].as<unsigned long long>() + (size_t)l * b.lane_cap, *k1 = b.sort_keys[1].as<unsigned long long>() + (size_t)l * b.lane_cap;
                        uint32_t *v0 = b.sort_vals[0].as<uint32_t>() + (size_t)l * b.lane_cap, *v1 = b.sort_vals[1].as<uint32_t>() + (size_t)l * b.lane_cap;
                        k_ray_keys<<<(nmax + 255) / 256, 256, 0, s2>>>(V, qi[l], half, nmax, lo2, inv, bits, k0, v0);
                        size_t tmp = b.sort_tmp_bytes;
                        GK_CUDA(cub::DeviceRadixSort::SortPairs(b.sort_tmp.p, tmp, k0, k1, v0, v1, (int)nmax, 0, 3 * bits + 5, s2));
                        Vt.order = v1;
                        c.launches += 2;
                    }
                    if (c.kernel_timing) cudaEventRecord(c.wev[l][3 * k], s2);
                    const uint32_t thr = tail_threshold;
                    if (c.trace_variant == 2) {
                        const uint32_t blocks = std::min<uint32_t>((bound[l] + 127) / 128, (uint32_t)c.sm_count * 8u);
                        k_trace_simple<<<blocks, 128, 0, s2>>>(S.ts, Vt, qi[l], half, thr);
                    } else if (c.trace_variant == 0) {
                        const uint32_t blocks = std::min<uint32_t>((bound[l] + 127) / 128, (uint32_t)c.sm_count * 8u);
                        if (c.node_format == 8) k_trace<0, 8><<<blocks, 128, 0, s2>>>(S.ts, Vt, qi[l], half, c.refill_below, thr);
                        else k_trace<0, 4><<<blocks, 128, 0, s2>>>(S.ts, Vt, qi[l], half, c.refill_below, thr);
                    } else if (c.trace_variant == 7 && c.node_format == 4) {
                        const uint32_t pb = std::min<uint32_t>((bound[l] + 4 * GK_WPOOL - 1) / (4 * GK_WPOOL), (uint32_t)c.sm_count * GK_WPOOL_MINB);
                        k_trace_wpool<4><<<pb, 128, 0, s2>>>(S.ts, Vt, qi[l], half, b.pool_stack.as<uint2>() + (size_t)l * b.pool_stack_stride, thr);
                    } else if (c.trace_variant == 6 && c.node_format == 4) {
                        const uint32_t pb = std::min<uint32_t>((bound[l] + GK_POOL - 1) / GK_POOL, (uint32_t)c.sm_count * GK_POOL_MINB);
                        k_trace_pool<4><<<pb, 128, 0, s2>>>(S.ts, Vt, qi[l], half, b.pool_stack.as<uint2>() + (size_t)l * b.pool_stack_stride, thr);
                    } else {
                        // rays per block the grid is sized for: with 128 every block of a wave that fits one residency round takes one
                        // batch and only drains it; a multiple leaves rays in the queue for the refills (GKCU_TRACE_RPB)
                        const uint32_t rpb = (uint32_t)c.trace_rpb;
                        const uint32_t tb = std::max<uint32_t>(1u, std::min<uint32_t>((bound[l] + rpb - 1) / rpb, (uint32_t)c.sm_count * (uint32_t)c.trace_grid_mult));
                        if (c.node_format == 8) k_trace<5, 8><<<tb, 128, 0, s2>>>(S.ts, Vt, qi[l], half, c.refill_below, thr);
                        else k_trace<5, 4><<<tb, 128, 0, s2>>>(S.ts, Vt, qi[l], half, c.refill_below, thr);
                    }
                    if (c.kernel_timing) cudaEventRecord(c.wev[l][3 * k + 1], s2);
                    const uint32_t sb = (bound[l] + 127) / 128;
                    if (c.shade_minb == 8) k_shade<8><<<sb, 128, 0, s2>>>(S, ubo, R, V, cfg, qi[l], ray_counters_of(l));
                    else if (c.shade_minb == 7) k_shade<7><<<sb, 128, 0, s2>>>(S, ubo, R, V, cfg, qi[l], ray_counters_of(l));
                    else if (c.shade_minb == 6) k_shade<6><<<sb, 128, 0, s2>>>(S, ubo, R, V, cfg, qi[l], ray_counters_of(l));
                    else if (c.shade_minb == 10) k_shade<10><<<sb, 128, 0, s2>>>(S, ubo, R, V, cfg, qi[l], ray_counters_of(l));
                    else k_shade<1><<<sb, 128, 0, s2>>>(S, ubo, R, V, cfg, qi[l], ray_counters_of(l));
                    if (c.kernel_timing) cudaEventRecord(c.wev[l][3 * k + 2], s2);
                    c.launches += 2;
                    qi[l] ^= 1;
                    ++iterations;
                }
                cudaEventRecord(c.pev[l][k], s2);  // once it has passed, ring word k holds the size this wave's k_trace stored
                ++launched[l];
                if (consume(l, launched[l] - seen[l] > ahead)) return 1;
                any = any || active[l];
            }
        }
        for (int l = 0; l < lanes; ++l) {  // the waves still in flight are timed once the frame has been synchronised (fetch_ray_counters)
            c.wave_seen[l] = seen[l];
            c.wave_launched[l] = launched[l];
        }
    } else {
        for (int l = 0; l < 4; ++l) c.wave_seen[l] = c.wave_launched[l] = 0;
    }
    for (int l = 0; l < lanes; ++l) {
        const uint32_t blocks = std::min<uint32_t>((pixels / lanes + 127) / 128, (uint32_t)c.sm_count * 16u);
        cudaEventRecord(c.lane_ev[l][0], ls[l]);
        if (p.tracer == 0 && c.tail_fused && c.node_format == 4) {
            // one warp per 32 paths up to what is resident at once; the warps refill from the queue beyond that
            const uint32_t fb = std::min<uint32_t>((std::max<uint32_t>(tail_threshold == 0xFFFFFFFFu ? pixels / lanes : tail_threshold, 128u) + 127) / 128, (uint32_t)c.sm_count * 4u);
            k_tail_fused<<<fb, 128, 0, ls[l]>>>(S, ubo, R, L.v[l], cfg, qi[l], ray_counters_of(l), 1);
        } else if (p.tracer == 0 && c.tail_segments > 0) {
            for (int sg = 0; sg <= c.tail_segments; ++sg)  // the last segment runs every remaining path to its end
                k_tail_seg<<<blocks, 128, 0, ls[l]>>>(S, ubo, R, L.v[l], cfg, qi[l], ray_counters_of(l), sg, sg < c.tail_segments ? c.tail_segment_rays[sg] : 0);
            c.launches += c.tail_segments;
        } else {
            const uint32_t* order = nullptr;
            if (p.tracer == 0 && c.tail_sort && ubo.NumberOfSamples > 1) {  // with one sample per pixel the paths end within a few rays of each other: nothing to order
                uint32_t* hist = b.tail_hist.as<uint32_t>() + 256 * l;
                uint32_t* ord = b.tail_order.as<uint32_t>() + (size_t)l * b.lane_cap;
                const uint32_t sb = std::max<uint32_t>(1u, std::min<uint32_t>((tail_threshold == 0xFFFFFFFFu ? pixels / lanes : tail_threshold) / 256u + 1u, (uint32_t)c.sm_count * 4u));
                GK_CUDA(cudaMemsetAsync(hist, 0, 256 * 4, ls[l]));
                k_tail_hist<<<sb, 256, 0, ls[l]>>>(L.v[l], qi[l], hist);
                k_tail_scan<<<1, 32, 0, ls[l]>>>(hist);
                k_tail_scatter<<<sb, 256, 0, ls[l]>>>(L.v[l], qi[l], hist, ord);
                c.launches += 3;
                order = ord;
            }
            k_tail<<<blocks, 128, 0, ls[l]>>>(S, ubo, R, L.v[l], cfg, qi[l], ray_counters_of(l), p.tracer == 0 ? 1 : 0, order, c.tail_spread, c.tail_solo);
        }
        cudaEventRecord(c.lane_ev[l][1], ls[l]);
        c.launches++;
        // per lane: primary / closest / any ray counters (3 x u64) and the wavefront statistics (rays, launches)
        if (publish_words(c, L.v[l].counts + 4, 64 + 64 * l + 4, 10, ls[l])) return 1;  // counts[4..9] ray counters, [12..13] wavefront statistics
#ifdef GK_TRAV_STATS
        if (publish_words(c, L.v[l].counts + 32, 64 + 64 * l + 32, 32, ls[l])) return 1;
#endif
        if (lanes > 1) {
            cudaEventRecord(c.lane_ev[l][2], ls[l]);
            cudaStreamWaitEvent(st, c.lane_ev[l][2], 0);  // the post passes on the context stream follow every lane
        }
    }
    c.kstats = ks;
    c.counters.wavefront_iterations = iterations;
    GK_CUDA(cudaGetLastError());
    return 0;
}

// valid after the stream has been synchronised
void fetch_ray_counters(Context& c) {
    if (!c.path || !c.path->host_counts) return;
    c.counters.primary_rays = c.counters.closest_hit_rays = c.counters.any_hit_rays = 0;
    c.kstats.trace_rays = 0;
    c.kstats.trace_launches = 0;
    c.kstats.tail_ms = 0;
    for (int l = 0; l < c.path->lanes; ++l) {
        const uint32_t* hc = c.path->host_counts + 64 * l;
        const unsigned long long* h = reinterpret_cast<const unsigned long long*>(hc + 4);
        c.counters.primary_rays += h[0];
        c.counters.closest_hit_rays += h[1];
        c.counters.any_hit_rays += h[2];
        c.kstats.trace_rays += hc[12];
        c.kstats.trace_launches += hc[13];
#ifdef GK_TRAV_STATS
        {
            const unsigned long long* g = reinterpret_cast<const unsigned long long*>(hc + 32);
            static const char* names[8] = {"tlas_node_steps", "blas_node_steps", "children_hit", "pushes", "culled_pops", "instance_entries", "leaf_steps", "tri_tests"};
            fprintf(stderr, "TRAV_STATS wavefront_rays=%u", hc[12]);
            for (int k = 0; k < 8; ++k) fprintf(stderr, " closest_%s=%llu", names[k], g[k]);
            for (int k = 0; k < 8; ++k) fprintf(stderr, " any_%s=%llu", names[k], g[8 + k]);
            fprintf(stderr, "\n");
        }
#endif
        float ms = 0;
        if (cudaEventElapsedTime(&ms, c.lane_ev[l][0], c.lane_ev[l][1]) == cudaSuccess) c.kstats.tail_ms += ms;
        for (; c.kernel_timing && c.wave_seen[l] < c.wave_launched[l]; ++c.wave_seen[l]) {  // waves the host had not looked at when it queued the tail
            const int k = c.wave_seen[l] % 8;
            float a = 0, bms = 0;
            if (cudaEventElapsedTime(&a, c.wev[l][3 * k], c.wev[l][3 * k + 1]) == cudaSuccess) c.kstats.trace_ms += a;
            if (cudaEventElapsedTime(&bms, c.wev[l][3 * k + 1], c.wev[l][3 * k + 2]) == cudaSuccess) c.kstats.shade_ms += bms;
        }
        c.wave_seen[l] = c.wave_launched[l];
    }
    c.kstats.shade_launches = c.kstats.trace_launches;
    float ms = 0;
    if (cudaEventElapsedTime(&ms, c.kev[0], c.kev[1]) == cudaSuccess) c.kstats.primary_ms = ms;
    cudaGetLastError();
    c.kstats.tail_rays = c.counters.closest_hit_rays + c.counters.any_hit_rays - c.kstats.trace_rays;
}

int run_resolve(Context& c, const GkUniformBufferObject& ubo, const GkcuFrameParams& p) {
    (void)ubo;
    (void)p;
    RTView R = make_rt_view(c);
    PathView V = make_path_view(c);
    uint32_t n = R.w * R.h;
    k_resolve<<<(n + 255) / 256, 256, 0, c.stream>>>(R, V.radiance, V.direct, R.moment, c.ubo.as<GkUniformBufferObject>());
    c.launches++;
    GK_CUDA(cudaGetLastError());
    return 0;
}

void* peer_local_buffer(Context& c, uint32_t what) {
    if (!c.path) return nullptr;
    switch (what) {
        case GKCU_PEER_RADIANCE: return c.path->radiance.p;
        case GKCU_PEER_DIRECT: return c.path->direct.p;
        case GKCU_PEER_SINGLE_DIFFUSE: return c.rt.single_diffuse;
        case GKCU_PEER_SINGLE_SPECULAR: return c.rt.single_specular;
        case GKCU_PEER_ACCUM_DIFFUSE_0: return c.rt.accum_diffuse[0];
        case GKCU_PEER_ACCUM_DIFFUSE_1: return c.rt.accum_diffuse[1];
        case GKCU_PEER_ACCUM_SPECULAR_0: return c.rt.accum_specular[0];
        case GKCU_PEER_ACCUM_SPECULAR_1: return c.rt.accum_specular[1];
        case GKCU_PEER_ACCUM_ALBEDO_0: return c.rt.accum_albedo[0];
        case GKCU_PEER_ACCUM_ALBEDO_1: return c.rt.accum_albedo[1];
        case GKCU_PEER_DENOISED: return c.rt.denoised;
        case GKCU_PEER_FLAGS:
            if (!c.peer_flags.p) {
                if (!c.peer_flags.reserve(32 * 8)) return nullptr;
                cudaMemset(c.peer_flags.p, 0, 32 * 8);
            }
            return c.peer_flags.p;
        default: return nullptr;
    }
}

int run_resolve_strip(Context& c, const GkUniformBufferObject& ubo, const GkcuFrameParams& p) {
    const int world = (int)std::max(1u, p.spp_world);
    if (world > GKCU_MAX_PEERS) { set_error("gkcu_render_resolve_strip: spp_world > 8"); return 1; }
    if ((int)p.spp_rank >= world) { set_error("gkcu_render_resolve_strip: spp_rank >= spp_world"); return 1; }
    PeerRadiance P{};
    for (int r = 0; r < world; ++r) {
        P.rad[r] = static_cast<const float4*>(c.peer[r][GKCU_PEER_RADIANCE]);
        if (!P.rad[r]) { set_error("gkcu_render_resolve_strip: radiance buffer of rank " + std::to_string(r) + " not imported"); return 1; }
    }
    void *d0 = c.peer[0][GKCU_PEER_DIRECT], *sd = c.peer[0][GKCU_PEER_SINGLE_DIFFUSE], *ss = c.peer[0][GKCU_PEER_SINGLE_SPECULAR];
    if (!d0 || !sd || !ss) { set_error("gkcu_render_resolve_strip: rank 0's direct / SINGLE_* buffers not imported"); return 1; }
    const uint32_t n = c.rt.w * c.rt.h;
    const uint32_t y0 = (uint32_t)((uint64_t)p.spp_rank * c.rt.h / world), y1 = (uint32_t)((uint64_t)(p.spp_rank + 1) * c.rt.h / world);
    const uint32_t first = y0 * c.rt.w, last = y1 * c.rt.w;
    // device-side barriers when every rank's flag block is known, otherwise the host brackets this call with its own
    bool flags = true;
    PeerFlags F{};
    for (int r = 0; r < world; ++r) {
        F.block[r] = static_cast<unsigned long long*>(c.peer[r][GKCU_PEER_FLAGS]);
        flags = flags && F.block[r];
    }
    unsigned long long* mine = flags ? F.block[p.spp_rank] : nullptr;
    if (flags) {
        ++c.peer_epoch;
        k_peer_signal<<<1, 32, 0, c.stream>>>(F, world, (int)p.spp_rank, 0, c.peer_epoch);   // my sums are complete
        k_peer_wait<<<1, 32, 0, c.stream>>>(mine, world, 0, c.peer_epoch);                    // ... and so are everybody's
        c.launches += 2;
    }
    if (last > first)
        k_resolve_strip<<<(last - first + 255) / 256, 256, 0, c.stream>>>(static_cast<uint2*>(sd), static_cast<uint2*>(ss), P, world, static_cast<const float4*>(d0),
                                                                       c.rt.moment, ubo, n, first, last);
    if (flags) {
        k_peer_signal<<<1, 32, 0, c.stream>>>(F, world, (int)p.spp_rank, 8, c.peer_epoch);   // my strip has landed; I no longer read anyone's sums
        k_peer_wait<<<1, 32, 0, c.stream>>>(mine, world, 8, c.peer_epoch);                    // all strips landed: rank 0 may post-process, everyone may start the next frame
        c.launches += 2;
    }
    c.launches++;
    GK_CUDA(cudaGetLastError());
    return 0;
}

void path_preload() {  // see post_preload (gk_post.cu)
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, k_resolve_strip);
    cudaFuncGetAttributes(&fa, k_resolve);
    cudaFuncGetAttributes(&fa, k_peer_signal);
    cudaFuncGetAttributes(&fa, k_peer_wait);
}

// The tail of a multi-GPU frame with nothing left to rank 0 alone, see gkcu_render_resolve_post_strip (include/gkcu.h).
int run_resolve_post_strip(Context& c, const GkUniformBufferObject& ubo, const GkcuFrameParams& p) {
    const int world = (int)std::max(1u, p.spp_world), rank = (int)p.spp_rank;
    if (world > GKCU_MAX_PEERS) { set_error("gkcu_render_resolve_post_strip: spp_world > 8"); return 1; }
    if (rank >= world) { set_error("gkcu_render_resolve_post_strip: spp_rank >= spp_world"); return 1; }
    if (c.rt.denoised_ext) { set_error("gkcu_render_resolve_post_strip: RT_DENOISED is bound to external memory; use gkcu_render_resolve_strip + gkcu_run_post"); return 1; }
    const int cur = c.rt.cur;
    PeerRadiance P{};
    PeerFlags F{};
    PostStrip S;
    for (int r = 0; r < world; ++r) {
        P.rad[r] = static_cast<const float4*>(c.peer[r][GKCU_PEER_RADIANCE]);
        F.block[r] = static_cast<unsigned long long*>(c.peer[r][GKCU_PEER_FLAGS]);
        void *ad = c.peer[r][GKCU_PEER_ACCUM_DIFFUSE_0 + cur], *as = c.peer[r][GKCU_PEER_ACCUM_SPECULAR_0 + cur], *aa = c.peer[r][GKCU_PEER_ACCUM_ALBEDO_0 + cur];
        if (!P.rad[r] || !F.block[r] || !ad || !as || !aa || !c.peer[r][GKCU_PEER_ACCUM_DIFFUSE_0 + (cur ^ 1)] || !c.peer[r][GKCU_PEER_ACCUM_SPECULAR_0 + (cur ^ 1)] ||
            !c.peer[r][GKCU_PEER_ACCUM_ALBEDO_0 + (cur ^ 1)]) {
            set_error("gkcu_render_resolve_post_strip: radiance / flags / ACCUM_* buffers of rank " + std::to_string(r) + " not imported");
            return 1;
        }
        if (r != rank) {
            S.mirror_diffuse[S.n_mirror] = static_cast<uint2*>(ad);
            S.mirror_specular[S.n_mirror] = static_cast<uint2*>(as);
            S.mirror_albedo[S.n_mirror] = static_cast<uint2*>(aa);
            ++S.n_mirror;
        }
    }
    void *d0 = c.peer[0][GKCU_PEER_DIRECT], *dn0 = c.peer[0][GKCU_PEER_DENOISED];
    if (!d0 || !dn0) { set_error("gkcu_render_resolve_post_strip: rank 0's direct / RT_DENOISED buffers not imported"); return 1; }
    const int h = (int)c.rt.h, w = (int)c.rt.w;
    const uint32_t n = (uint32_t)w * (uint32_t)h;
    S.own_y0 = (int)((uint64_t)rank * h / world);
    S.own_y1 = (int)((uint64_t)(rank + 1) * h / world);
    S.y_begin = std::max(0, S.own_y0 - 5);   // the compose pass reads ACCUM_DIFFUSE 5 rows either side (Process.DenoiseJBF.comp.slang:133)
    S.y_end = std::min(h, S.own_y1 + 5);
    const int ry0 = std::max(0, S.y_begin - 2), ry1 = std::min(h, S.y_end + 2);  // the reproject pass reads SINGLE_* 2 rows either side (:120-139)
    S.denoised = rank == 0 ? nullptr : static_cast<uint2*>(dn0);
    cudaStream_t st = c.stream;
    if (rank == 0 && c.copy_pending) {  // the asynchronous present of the previous frame still reads RT_DENOISED, which the other ranks are about to write
        cudaStreamWaitEvent(st, c.copy_done, 0);
        c.copy_pending = false;
    }
    unsigned long long* mine = F.block[rank];
    ++c.peer_epoch;
    k_peer_signal<<<1, 32, 0, st>>>(F, world, rank, 0, c.peer_epoch);   // my sums are complete (and rank 0's image is free)
    k_peer_wait<<<1, 32, 0, st>>>(mine, world, 0, c.peer_epoch);        // ... and so are everybody's
    const uint32_t first = (uint32_t)ry0 * (uint32_t)w, last = (uint32_t)ry1 * (uint32_t)w;
    if (last > first)
        k_resolve_strip<<<(last - first + 255) / 256, 256, 0, st>>>(c.rt.single_diffuse, c.rt.single_specular, P, world, static_cast<const float4*>(d0), c.rt.moment, ubo, n, first, last);
    c.launches += 3;
    if (run_reproject(c, p, &S)) return 1;
    if (run_denoise(c, ubo, &S)) return 1;
    k_peer_signal<<<1, 32, 0, st>>>(F, world, rank, 8, c.peer_epoch);   // my strip of RT_DENOISED and of everyone's history has landed; I no longer read anyone's sums
    k_peer_wait<<<1, 32, 0, st>>>(mine, world, 8, c.peer_epoch);        // all strips landed: rank 0 may present, everyone may start the next frame
    c.launches += 2;
    GK_CUDA(cudaGetLastError());
    return 0;
}

int radiance_buffer(Context& c, void** ptr, size_t* bytes) {
    if (!c.path) { set_error("gkcu_radiance_buffer: call gkcu_resize first"); return 1; }
    *ptr = c.path->radiance.p;
    *bytes = 2 * c.path->pixels * 16;
    return 0;
}

int run_trace_rays(Context& c, uint32_t count, bool any_hit) {
    k_trace_rays<<<(count + 127) / 128, 128, 0, c.stream>>>(make_trace_scene(c), c.trace_rays.as<GkcuRay>(), count, c.trace_hits.as<GkcuHit>(), any_hit ? 1 : 0);
    c.launches++;
    GK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace gkcu
