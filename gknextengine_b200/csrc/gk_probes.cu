// gk_probes.cu — ambient-cube probe grid: voxelisation, page index and the light bake.
//
// The path tracer ends long paths by looking up this grid (interpolateAmbientCubesV2, gk_shading.cuh); the reference fills it
// from two places, both moved to the GPU here on top of the same BVH and tracer as the rt pass:
//   * voxels + page index: CPU, FCPUProbeBaker::ProcessCube / VoxelizeCube and FCPUPageIndex::UpdateData
//     (src/Assets/CPUAccelerationStructure.cpp:85-159, 313-328, 499-528) over tinybvh -> k_voxelize, k_page_index
//   * cubes: GPU, Bake.HwAmbientCube.comp.slang -> FGpuProbeGenerator::RenderAmbientCubeOnly / FaceTask
//     (assets/shaders/common/AmbientCube.slang:471-553, 673-692), 1/120 of the grid per frame -> k_bake
#include "gk_context.h"
#include "gk_shading.cuh"

using namespace gkd;

namespace gkcu {

namespace {

constexpr float kCubeUnit = 0.25f;                 // CUBE_UNIT, AmbientCube.slang:14
constexpr float kFastMaxTraceDistance = 20.0f;     // FAST_MAX_TRACE_DISTANCE, Shading.slang:16
constexpr uint32_t kGridCount = GK_CUBE_SIZE_XY * GK_CUBE_SIZE_XY * GK_CUBE_SIZE_Z;

GK_DEV float3 probe_origin(uint32_t gIdx) {  // Bake.HwAmbientCube.comp.slang:12-17, CUBE_OFFSET AmbientCube.slang:15
    const uint32_t y = gIdx / (GK_CUBE_SIZE_XY * GK_CUBE_SIZE_XY);
    const uint32_t z = (gIdx - y * GK_CUBE_SIZE_XY * GK_CUBE_SIZE_XY) / GK_CUBE_SIZE_XY;
    const uint32_t x = gIdx - y * GK_CUBE_SIZE_XY * GK_CUBE_SIZE_XY - z * GK_CUBE_SIZE_XY;
    const float3 CUBE_OFFSET = f3(-GK_CUBE_SIZE_XY / 2, -1.375f, -GK_CUBE_SIZE_XY / 2) * kCubeUnit;
    return f3((float)x, (float)y, (float)z) * kCubeUnit + CUBE_OFFSET;
}

// ------------------------------------------------------------------------------------------------ voxelisation
// TraceRay of the CPU baker (CPUAccelerationStructure.cpp:42-66): tinybvh::Ray normalises the direction and has no tmin;
// a hit needs t < dist. The normal is the geometric face normal taken through the node's world matrix (not renormalised),
// the material comes from vertex 0's slot. Geometry is the fp16 vertex buffer the rt pass traces (the CPU baker keeps fp32).
struct VoxHit { float t; float3 nrm; uint32_t mat; };
GK_DEV bool vox_trace(const SceneView& S, float3 origin, float3 dir, float dist, VoxHit& out) {
    const float3 d = normalize(dir);
    Hit h;
    trace_ray<false>(S.ts, origin, d, 0.0f, dist, h);
    if (h.inst == GK_NO_HIT) return false;
    const GkNodeProxy* node = S.nodes + h.inst;
    const GkModelData off = S.offsets[node->modelId];
    const uint32_t i0 = off.vertexOffset + S.indices[off.indexOffset + h.prim * 3], i1 = off.vertexOffset + S.indices[off.indexOffset + h.prim * 3 + 1],
                   i2 = off.vertexOffset + S.indices[off.indexOffset + h.prim * 3 + 2];
    const VtxLite v0 = unpack_vertex(S, i0), v1 = unpack_vertex(S, i1), v2 = unpack_vertex(S, i2);
    const float3 n = normalize(cross(v1.pos - v0.pos, v2.pos - v1.pos));  // InitBVH, :190-193
    const GkMat4& M = node->worldTS;
    out.nrm = f3(M.m[0][0] * n.x + M.m[1][0] * n.y + M.m[2][0] * n.z, M.m[0][1] * n.x + M.m[1][1] * n.y + M.m[2][1] * n.z,
                 M.m[0][2] * n.x + M.m[1][2] * n.y + M.m[2][2] * n.z);
    out.t = h.t;
    out.mat = fetch_material_id(node, v0.matSlot);
    return true;
}

GK_DEV void inside_geometry(const SceneView& S, float3 origin, float3 rayDir, uint32_t& matId, float& distance) {  // :85-110
    VoxHit h;
    if (vox_trace(S, origin, rayDir, kCubeUnit * 64, h)) {
        distance = h.t;
        if (distance <= kCubeUnit) {
            matId = h.mat;
            const GkMaterial m = get_material(S, h.mat);
            if (dot(h.nrm, rayDir) > 0.0f || m.MaterialModel == GK_MATERIAL_DIFFUSE_LIGHT) distance = 0.0f;
        }
    }
}
GK_DEV float detect_distance(const SceneView& S, float3 origin, float3 rayDir) {  // :72-83
    VoxHit h;
    return vox_trace(S, origin, rayDir, kCubeUnit * 64, h) ? h.t : 255.0f;
}

__global__ void __launch_bounds__(128) k_voxelize(SceneView S, ProbeRec* __restrict__ probes, ProbeMeta* __restrict__ meta) {  // VoxelizeCube, :112-159
    const uint32_t gIdx = blockIdx.x * blockDim.x + threadIdx.x;
    if (gIdx >= kGridCount) return;
    const float3 origin = probe_origin(gIdx);
    GkVoxelData cube;
    cube.age = 0;
    cube.matId = 0;
    float distPY = 255.0f, distNY = 255.0f, distPX = 255.0f, distNX = 255.0f, distPZ = 255.0f, distNZ = 255.0f;
    if (S.ts.tlas_nodes) {
        inside_geometry(S, origin, f3(0, 1, 0), cube.matId, distPY);
        inside_geometry(S, origin, f3(0, -1, 0), cube.matId, distNY);
        inside_geometry(S, origin, f3(1, 0, 0), cube.matId, distPX);
        inside_geometry(S, origin, f3(-1, 0, 0), cube.matId, distNX);
        inside_geometry(S, origin, f3(0, 0, 1), cube.matId, distPZ);
        inside_geometry(S, origin, f3(0, 0, -1), cube.matId, distNZ);
    }
    float minDist = fminf(fminf(fminf(distPY, distNY), fminf(distPX, distNX)), fminf(distPZ, distNZ));
    if (minDist > 254.0f && S.ts.tlas_nodes) {  // the reference's list repeats (-1,1,1) and (-1,1,-1); kept, min() is idempotent
        minDist = fminf(minDist, detect_distance(S, origin, f3(1, 1, 1)));
        minDist = fminf(minDist, detect_distance(S, origin, f3(-1, 1, 1)));
        minDist = fminf(minDist, detect_distance(S, origin, f3(-1, -1, 1)));
        minDist = fminf(minDist, detect_distance(S, origin, f3(1, 1, -1)));
        minDist = fminf(minDist, detect_distance(S, origin, f3(-1, 1, -1)));
        minDist = fminf(minDist, detect_distance(S, origin, f3(-1, -1, -1)));
    }
    distPY = clampf(distPY / kCubeUnit, 0.0f, 1.0f); distNY = clampf(distNY / kCubeUnit, 0.0f, 1.0f);
    distPX = clampf(distPX / kCubeUnit, 0.0f, 1.0f); distNX = clampf(distNX / kCubeUnit, 0.0f, 1.0f);
    distPZ = clampf(distPZ / kCubeUnit, 0.0f, 1.0f); distNZ = clampf(distNZ / kCubeUnit, 0.0f, 1.0f);
    const float inside = distPY * distNY * distPX * distNX * distPZ * distNZ;
    cube.distanceToSolid_gg_z01 = ((uint32_t)(minDist / kCubeUnit) & 0xFF) | (((uint32_t)(inside * 255.0f) & 0xFF) << 8) |
                                  (((uint32_t)(distPZ * 255.0f) & 0xFF) << 16) | (((uint32_t)(distNZ * 255.0f) & 0xFF) << 24);
    cube.distanceToSolid_x01_y01 = ((uint32_t)(distPX * 255.0f) & 0xFF) | (((uint32_t)(distNX * 255.0f) & 0xFF) << 8) |
                                   (((uint32_t)(distPY * 255.0f) & 0xFF) << 16) | (((uint32_t)(distNY * 255.0f) & 0xFF) << 24);
    ProbeRec rec = {};  // colours cleared: a full (non-incremental) pass starts the bake over
    rec.dist0 = cube.distanceToSolid_gg_z01;
    rec.dist1 = cube.distanceToSolid_x01_y01;
    probes[gIdx] = rec;
    meta[gIdx] = ProbeMeta{cube.matId, cube.age};
}

// FCPUPageIndex::UpdateData / GetPage, :499-545: voxels with a material, counted per 16 m page of the 64 x 64 index.
__global__ void k_page_index(const ProbeMeta* __restrict__ meta, GkPageIndex* __restrict__ pages) {
    const uint32_t gIdx = blockIdx.x * blockDim.x + threadIdx.x;
    if (gIdx >= kGridCount) return;
    if (meta[gIdx].matId == 0) return;
    const float3 rel = probe_origin(gIdx) - f3(-512.0f, 0.0f, -512.0f);
    const int px = min(max((int)(rel.x / 16.0f), 0), GK_PAGE_COUNT - 1), pz = min(max((int)(rel.z / 16.0f), 0), GK_PAGE_COUNT - 1);
    atomicAdd(&pages[pz * GK_PAGE_COUNT + px].voxelCount, 1u);
}

// ------------------------------------------------------------------------------------------------ bake
__constant__ float2 c_grid3x3[9] = {{-0.667f, -0.667f}, {0.0f, -0.667f}, {0.667f, -0.667f}, {-0.667f, 0.0f}, {0.0f, 0.0f},
                                    {0.667f, 0.0f},     {-0.667f, 0.667f}, {0.0f, 0.667f},  {0.667f, 0.667f}};
__constant__ float2 c_grid4x4[16] = {{-0.75f, -0.75f}, {-0.25f, -0.75f}, {0.25f, -0.75f}, {0.75f, -0.75f}, {-0.75f, -0.25f}, {-0.25f, -0.25f},
                                     {0.25f, -0.25f},  {0.75f, -0.25f},  {-0.75f, 0.25f}, {-0.25f, 0.25f}, {0.25f, 0.25f},   {0.75f, 0.25f},
                                     {-0.75f, 0.75f},  {-0.25f, 0.75f},  {0.25f, 0.75f},  {0.75f, 0.75f}};

// One block per probe, one thread per (face, ray): 6 faces x FACE_TRACING_COUNT = 96 rays in flight per probe. The per-ray
// terms go through shared memory and are summed by one thread per face in ray order, so the sums round exactly like the
// reference's sequential loop (AmbientCube.slang:494-518). Faces in the call order of RenderAmbientCubeOnly (:685-690).
__global__ void __launch_bounds__(96) k_bake(SceneView S_, const __grid_constant__ GkUniformBufferObject cam, SnapshotCubes F, ProbeRec* __restrict__ probes,
                                             ProbeMeta* __restrict__ meta, uint32_t first) {
    SceneView S = S_;
    S.cam = &cam;
    const uint32_t gIdx = first + blockIdx.x;
    if ((probes[gIdx].dist0 & 0xFFu) >= 8u) return;  // only probes near a surface, :679
    __shared__ float4 s_term[6][16];
    __shared__ uint32_t s_hit[6][16];
    __shared__ uint32_t s_skyvis[6], s_sunvis;
    const int face = threadIdx.x >> 4, i = threadIdx.x & 15;
    const float3 basis = face == 0 ? f3(0, 1, 0) : face == 1 ? f3(0, -1, 0) : face == 2 ? f3(1, 0, 0) : face == 3 ? f3(-1, 0, 0) : face == 4 ? f3(0, 0, 1) : f3(0, 0, -1);
    const uint32_t iterate = meta[gIdx].age;
    const float3 origin = probe_origin(gIdx) + basis * kCubeUnit * 0.25f;
    {   // FaceTask, :494-518
        const float2 offset = c_grid3x3[iterate % 9u] * 0.25f;
        const float2 g = c_grid4x4[i] + offset;
        const float3 rayDir = AlignWithNormal(normalize(f3(g.x, g.y, 1.0f)), basis);
        Hit h;
        trace_ray<false>(S.ts, origin, rayDir, GK_EPS, kFastMaxTraceDistance, h);
        float4 term;
        if (h.inst != GK_NO_HIT) {
            const HitVertex hv = resolve_hit(S, origin, rayDir, h);
            const GkMaterial m = get_material(S, hv.mat);
            float4 outAlbedo = f4(m.Diffuse.x, m.Diffuse.y, m.Diffuse.z, m.Diffuse.w);
            if (m.DiffuseTextureId >= 0) outAlbedo = outAlbedo * sample_texture(S, (uint32_t)m.DiffuseTextureId, hv.uv);
            term = outAlbedo * interpolateAmbientCubesV2<DIAmbientCubeSampler>(S, F, hv.pos, hv.nrm) * 1.25f;  // "magic bounce twice"
        } else {
            term = SampleIBLRough(S, cam.SkyIdx, rayDir, cam.SkyRotation) * (cam.HasSky ? cam.SkyIntensity : 0.0f);
        }
        s_term[face][i] = term;
        s_hit[face][i] = h.inst != GK_NO_HIT ? 1u : 0u;
    }
    __syncthreads();
    if (i == 0) {
        float4 directColor = f4(0, 0, 0, 0), bounceColor = f4(0, 0, 0, 0);
        float skyVisibility = 0.0f;
        for (int k = 0; k < 16; ++k) {
            if (s_hit[face][k]) bounceColor = bounceColor + s_term[face][k];
            else { directColor = directColor + s_term[face][k]; skyVisibility += 1.0f; }
        }
        directColor = directColor / 16.0f;
        bounceColor = bounceColor / 16.0f;
        if (cam.LightCount > 0 && S.lights) {  // one parametric light, :524-535
            const GkLightObject light = S.lights[0];
            const GkMaterial lm = get_material(S, light.lightMatIdx);
            const float4 lightPower = f4(lm.Diffuse.x, lm.Diffuse.y, lm.Diffuse.z, lm.Diffuse.w);
            const float3 lightPos = lerp3(f3(light.p1.x, light.p1.y, light.p1.z), f3(light.p3.x, light.p3.y, light.p3.z), 0.5f);
            const float3 dir = lightPos - origin;
            const float len = length(dir);
            Hit sh;  // TraceSegment(origin, lightPos, CUBE_UNIT * 0.5), Shading.slang:759-783
            trace_ray<true>(S.ts, origin, f3(dir.x / len, dir.y / len, dir.z / len), kCubeUnit * 0.5f, len - kCubeUnit * 0.5f, sh);
            const float lightAtten = sh.inst != GK_NO_HIT ? 0.0f : 1.0f;
            const float3 lightDir = normalize(lightPos - origin);
            const float ndotl = clampf(dot(basis, lightDir), 0.0f, 1.0f);
            const float distance = length(lightPos - origin);
            const float attenuation = ndotl * light.normal_area.w / (distance * distance * 3.14159f);
            directColor = directColor + lightPower * attenuation * lightAtten;
        }
        uint32_t sunVis = 0;
        if (cam.HasSun) {  // :538-545
            const float3 sunDir = f3(cam.SunDirection.x, cam.SunDirection.y, cam.SunDirection.z);
            Hit sh;  // TraceOcclusion, Shading.slang:737-757
            trace_ray<true>(S.ts, origin, sunDir, GK_EPS, GK_PT_MAX_TRACE_DISTANCE, sh);
            const float sunAtten = sh.inst != GK_NO_HIT ? 0.0f : 1.0f;
            const float ndotl = clampf(dot(basis, sunDir), 0.0f, 1.0f);
            sunVis = sunAtten > 0.0f ? 1u : 0u;
            const float4 sunColor = f4(cam.SunColor.x, cam.SunColor.y, cam.SunColor.z, cam.SunColor.w);
            directColor = directColor + sunColor * sunAtten * ndotl * 1.0f * 0.25f;
        }
        // temporal accumulation, :548-551. The stored cube is read from the snapshot (== the live value: only this block writes it).
        const GkAmbientCube old = F.fetch((int)gIdx);
        const uint32_t s0 = old.skyVisibility_pznzpyny, s1 = old.skyVisibility_pxnxs0s1;
        // byte of this face: +Y s0.z, -Y s0.w, +X s1.x, -X s1.y, +Z s0.x, -Z s0.y (:685-690)
        const uint32_t oldVis = face == 0 ? (s0 >> 16) & 0xFF : face == 1 ? (s0 >> 24) & 0xFF : face == 2 ? s1 & 0xFF : face == 3 ? (s1 >> 8) & 0xFF
                                : face == 4 ? s0 & 0xFF : (s0 >> 8) & 0xFF;
        const float currWeight = 0.125f;
        const float a = (float)oldVis, b = 255.0f * skyVisibility / 16.0f;
        s_skyvis[face] = (uint32_t)(a + (b - a) * currWeight);
        if (face == 0) s_sunvis = cam.HasSun ? sunVis : ((s1 >> 16) & 0xFF);  // SunVisOut is only assigned under HasSun; face +Y owns the byte
        const uint32_t oldD = face == 0 ? old.PosY_D : face == 1 ? old.NegY_D : face == 2 ? old.PosX_D : face == 3 ? old.NegX_D : face == 4 ? old.PosZ_D : old.NegZ_D;
        const uint32_t oldI = face == 0 ? old.PosY : face == 1 ? old.NegY : face == 2 ? old.PosX : face == 3 ? old.NegX : face == 4 ? old.PosZ : old.NegZ;
        const uint32_t newD = LerpPackedColorAlt(oldD, directColor, currWeight), newI = LerpPackedColorAlt(oldI, bounceColor, currWeight);
        // GkAmbientCube word order: PosZ NegZ PosY NegY PosX NegX, then the same six with _D
        const int w = face == 0 ? 2 : face == 1 ? 3 : face == 2 ? 4 : face == 3 ? 5 : face == 4 ? 0 : 1;
        probes[gIdx].cube[6 + w] = newD;
        probes[gIdx].cube[w] = newI;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const GkAmbientCube old = F.fetch((int)gIdx);
        probes[gIdx].cube[12] = (s_skyvis[4] & 0xFF) | ((s_skyvis[5] & 0xFF) << 8) | ((s_skyvis[0] & 0xFF) << 16) | ((s_skyvis[1] & 0xFF) << 24);
        probes[gIdx].cube[13] = (s_skyvis[2] & 0xFF) | ((s_skyvis[3] & 0xFF) << 8) | ((s_sunvis & 0xFF) << 16) | (old.skyVisibility_pxnxs0s1 & 0xFF000000u);
        meta[gIdx].age = iterate + 1;
    }
}

}  // namespace

int probes_voxelize(Context& c) {
    const size_t n = kGridCount;
    if (!c.probes.reserve(n * sizeof(ProbeRec)) || !c.probe_meta.reserve(n * sizeof(ProbeMeta)) ||
        !c.pages.reserve((size_t)GK_PAGE_COUNT * GK_PAGE_COUNT * sizeof(GkPageIndex)))
        return 1;
    SceneView S = make_scene_view(c);
    cudaStream_t st = c.stream;
    k_voxelize<<<(kGridCount + 127) / 128, 128, 0, st>>>(S, c.probes.as<ProbeRec>(), c.probe_meta.as<ProbeMeta>());
    GK_CUDA(cudaMemsetAsync(c.pages.p, 0, (size_t)GK_PAGE_COUNT * GK_PAGE_COUNT * sizeof(GkPageIndex), st));
    k_page_index<<<(kGridCount + 255) / 256, 256, 0, st>>>(c.probe_meta.as<ProbeMeta>(), c.pages.as<GkPageIndex>());
    GK_CUDA(cudaGetLastError());
    c.has_probes = true;
    c.launches += 2;
    return 0;
}

int probes_bake(Context& c, const GkUniformBufferObject& ubo, uint32_t first, uint32_t count) {
    if (!c.has_probes) { set_error("gkcu_probes_bake: no probe state (call gkcu_probes_voxelize or gkcu_set_probes)"); return 1; }
    if ((uint64_t)first + count > kGridCount) { set_error("gkcu_probes_bake: range outside the 192 x 48 x 192 grid"); return 1; }
    if (count == 0) return 0;
    if (!c.cube_snapshot.reserve((size_t)count * sizeof(ProbeRec))) return 1;
    cudaStream_t st = c.stream;
    GK_CUDA(cudaMemcpyAsync(c.cube_snapshot.p, c.probes.as<ProbeRec>() + first, (size_t)count * sizeof(ProbeRec), cudaMemcpyDeviceToDevice, st));
    SceneView S = make_scene_view(c);
    SnapshotCubes F{c.probes.as<ProbeRec>(), c.cube_snapshot.as<ProbeRec>(), first, count};
    k_bake<<<count, 96, 0, st>>>(S, ubo, F, c.probes.as<ProbeRec>(), c.probe_meta.as<ProbeMeta>(), first);
    GK_CUDA(cudaGetLastError());
    c.launches += 1;
    return 0;
}

}  // namespace gkcu
