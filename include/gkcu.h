/*
 * gkcu.h — C ABI of libgknext_cuda.so, the B200-native (sm_100a) CUDA path-tracing
 * renderer that stands in for gkNextEngine's Vulkan RayTracing renderer.
 *
 * Everything here is plain C: POD structs, raw host pointers, sizes and int return
 * codes (0 = ok, anything else = error; text via gkcu_last_error()). No exceptions,
 * no torch/CUDA types cross this boundary. One render thread per context.
 *
 * The POD structs are byte-exact mirrors of the reference's shared C++/Slang header
 * (reference: assets/shaders/common/BasicTypes.slang:22-291, included into C++ by
 * src/Assets/UniformBuffer.hpp:30-48 with bool -> uint32_t, half4 -> uint16). The
 * static_asserts at the bottom pin sizeof/offsetof to the values measured from the
 * reference header; oracle/ref_layout_check.cpp re-checks them
 * against the reference header itself when /root/reference is mounted.
 *
 * Each entry point cites the reference interface it replaces.
 */
#ifndef GKCU_H
#define GKCU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ POD types */

typedef struct GkVec4 { float x, y, z, w; } GkVec4;
/* glm column-major: m[col][row]; shaders use mul(M, v) == M * v. */
typedef struct GkMat4 { float m[4][4]; } GkMat4;

/* BasicTypes.slang:22-84 */
typedef struct GkUniformBufferObject {
    GkMat4 ModelView;
    GkMat4 Projection;
    GkMat4 ModelViewInverse;
    GkMat4 ProjectionInverse;
    GkMat4 ViewProjection;
    GkMat4 PrevViewProjection;
    GkMat4 ViewProjectionUnJit;
    GkMat4 PrevViewProjectionUnJit;

    GkVec4 ViewportRect;
    GkVec4 SunDirection;
    GkVec4 SunColor;
    GkVec4 BackGroundColor;

    GkMat4 SunViewProjection;

    float Aperture;
    float FocusDistance;
    float SkyRotation;
    float HeatmapScale;

    float PaperWhiteNit;
    float SkyIntensity;
    uint32_t SkyIdx;
    uint32_t TotalFrames;

    uint32_t MaxNumberOfBounces;
    uint32_t NumberOfSamples;
    uint32_t NumberOfBounces;
    uint32_t RandomSeed;

    uint32_t LightCount;
    uint32_t HasSky;
    uint32_t ShowHeatmap;
    uint32_t UseCheckerBoard;

    uint32_t TemporalFrames;
    uint32_t HasSun;
    uint32_t HDR;
    uint32_t AdaptiveSample;

    float AdaptiveVariance;
    uint32_t AdaptiveSteps;
    uint32_t TAA;
    uint32_t SelectedId;

    uint32_t ShowEdge;
    uint32_t ProgressiveRender;
    float BFSigma;
    float BFSigmaLum;

    float BFSigmaNormal;
    uint32_t BFSize;

    uint32_t FastGather;

    uint32_t FastInterpole;
    uint32_t DebugDraw_Lighting;
    uint32_t DisableSpatialReuse;
    uint32_t SuperResolution;
    uint32_t _pad0;
} GkUniformBufferObject;

/* BasicTypes.slang:86-95 */
typedef struct GkNodeProxy {
    uint32_t instanceId;
    uint32_t modelId;   /* model * 10 + section (Scene.cpp:470) */
    uint32_t visible;
    uint32_t nort;      /* section > 0 -> not in the TLAS (RayTraceBaseRenderer.cpp:178) */
    GkMat4 worldTS;
    GkMat4 combinedPrevTS;
    uint32_t matId[16];
} GkNodeProxy;

/* BasicTypes.slang:97-111 */
typedef struct GkModelData {
    uint32_t indexOffset;
    uint32_t indexCount;
    uint32_t vertexOffset;
    uint32_t vertexCount;
    GkVec4 localAabbMin;
    GkVec4 localAabbMax;
    uint32_t modelType;
    uint32_t voxelDataIdx;
    uint32_t reorderOffset;
    uint32_t reserved2;
} GkModelData;

/* BasicTypes.slang:116-138 */
typedef struct GkAmbientCube {
    uint32_t PosZ, NegZ, PosY, NegY, PosX, NegX;
    uint32_t PosZ_D, NegZ_D, PosY_D, NegY_D, PosX_D, NegX_D;
    uint32_t skyVisibility_pznzpyny;
    uint32_t skyVisibility_pxnxs0s1;
} GkAmbientCube;

/* BasicTypes.slang:145-151 */
typedef struct GkVoxelData {
    uint32_t matId;
    uint32_t age;
    uint32_t distanceToSolid_gg_z01;
    uint32_t distanceToSolid_x01_y01;
} GkVoxelData;

/* BasicTypes.slang:156-162: one entry of the 64 x 64 page index over the probe grid */
typedef struct GkPageIndex {
    uint32_t pageId;
    uint32_t voxelCount;
    uint32_t voxelDataIdx;
    uint32_t probeDataIdx;
} GkPageIndex;

/* BasicTypes.slang:164-169: 9-coefficient SH of a sky HDR, one row per colour channel */
typedef struct GkSphericalHarmonics {
    float coefficients[3][9];
    float padding;
} GkSphericalHarmonics;

/* BasicTypes.slang:171-183: quad light (corners p0, p1, p3; area in normal_area.w) */
typedef struct GkLightObject {
    GkVec4 p0, p1, p3, normal_area;
    uint32_t lightMatIdx, reserved1, reserved2, reserved3;
} GkLightObject;

/* BasicTypes.slang:218-236 / src/Assets/Material.hpp:8-74 */
typedef struct GkMaterial {
    GkVec4 Diffuse;
    int32_t DiffuseTextureId;
    int32_t MRATextureId;
    int32_t NormalTextureId;
    float Fuzziness;
    float RefractionIndex;
    uint32_t MaterialModel;
    float Metalness;
    float RefractionIndex2;
    float NormalTextureScale;
    float Reserverd2;
    float Reserverd3;
    float Reserverd4;
} GkMaterial;

enum { /* BasicTypes.slang:183-188 */
    GK_MATERIAL_LAMBERTIAN = 0,
    GK_MATERIAL_METALLIC = 1,
    GK_MATERIAL_DIELECTRIC = 2,
    GK_MATERIAL_ISOTROPIC = 3,
    GK_MATERIAL_DIFFUSE_LIGHT = 4,
    GK_MATERIAL_MIXTURE = 5
};

/* BasicTypes.slang:249-254 / src/Assets/Vertex.hpp:32-47: 12 x 16-bit.
 * pos.xyz, u, nrm.xyz, v, tan.xyz as IEEE half bits (glm::detail::toFloat16 rounding),
 * tangentw = (tanSign > 0 ? 2 : 0) << 8 | materialSlot (Vertex.hpp:97). */
typedef struct GkGPUVertex {
    uint16_t posx, posy, posz, texcoordx;
    uint16_t normalx, normaly, normalz, texcoordy;
    uint16_t tangentx, tangenty, tangentz, tangentw;
} GkGPUVertex;

/* Storage-image indices: assets/shaders/common/BindlessTexture.slang:8-40. */
enum {
    GK_RT_ACCUMLATE_DIFFUSE = 0,
    GK_RT_SINGLE_DIFFUSE = 1,
    GK_RT_MINIGBUFFER = 2,
    GK_RT_OBJEDCTID_0 = 3,
    GK_RT_OBJEDCTID_1 = 4,
    GK_RT_MOTIONVECTOR = 5,
    GK_RT_ALBEDO = 6,
    GK_RT_NORMAL = 7,
    GK_RT_SHADER_TIMER = 8,
    GK_RT_DENOISED = 9,
    GK_RT_PREV_DEPTHBUFFER = 10,
    GK_RT_ACCUMLATE_SPECULAR = 11,
    GK_RT_SINGLE_SPECULAR = 12,
    GK_RT_ACCUMLATE_ALBEDO = 13,
    GK_RT_MOTIONMOMENT = 14,
    GK_RT_SINGLE_PREV_DIFFUSE = 20,
    GK_RT_SINGLE_PREV_SPECULAR = 21,
    GK_RT_SINGLE_PREV_ALBEDO = 22,
    /* CUDA-side extension (no reference counterpart): per-pixel primitive index of the
     * primary hit (R32_UINT, 0xFFFFFFFF on miss), used for the id-parity check. */
    GK_RT_EXT_PRIMITIVEID = 64
};

#define GK_OBJECTID_MISS 65535u      /* Shading.slang:1199 */
#define GK_MAX_MATERIALS 8192        /* BasicTypes.slang:11 */

/* Probe grid: assets/shaders/common/AmbientCube.slang:11-15. */
#define GK_CUBE_SIZE_XY 192
#define GK_CUBE_SIZE_Z 48
#define GK_PAGE_COUNT 64 /* ACGI_PAGE_COUNT, src/Assets/UniformBuffer.hpp:19-21: 64 x 64 pages of 16 m from (-512, 0, -512) */

/* ------------------------------------------------------------- scene contract */

/* Outputs of Scene::RebuildMeshBuffer (reference: src/Assets/Scene.cpp:102-289):
 * the global vertex/index/offset arrays the shaders read through GPUScene
 * (BasicTypes.slang:261-291). All pointers are HOST memory; the call copies. */
typedef struct GkcuSceneDesc {
    const GkGPUVertex* vertices;      /* GPUScene.Vertices */
    uint32_t vertex_count;
    const uint16_t* simple_vertices;  /* GPUScene.VerticesSimple: half4 (x,y,z,x) per vertex, the BLAS positions
                                         (BottomLevelGeometry.cpp:21-24) */
    const uint32_t* indices;          /* GPUScene.Indices (= PrimAddress: model-local vertex ids) */
    uint32_t index_count;
    const GkModelData* offsets;       /* GPUScene.Offsets: 10 sections per model */
    uint32_t offset_count;            /* = 10 * model_count */
    uint32_t model_count;
} GkcuSceneDesc;

typedef struct GkcuFrameParams {
    uint32_t progressive;       /* ReProject push constant ProgressiveRender (PathTracingRenderer.cpp:127) */
    int32_t temporal_frames;    /* ReProject push constant TemporalFrames (user setting, not ubo.TemporalFrames) */
    uint32_t fast_reproject;    /* push constant FastReproject (always 0 in the reference) */
    /* Multi-GPU sample split (no reference counterpart). world = 1 reproduces the reference.
     * Rank r of `spp_world` traces samples [r*S/world, (r+1)*S/world) with frame seed TotalFrames*world + r. */
    uint32_t spp_rank;
    uint32_t spp_world;
    uint32_t tracer;            /* 0 = wavefront (default), 1 = megakernel (one thread per pixel) */
    uint32_t skip_post;         /* 1 = stop after the rt pass (radiance sums left un-normalised for a reduce) */
    uint32_t primary_visibility; /* 0 = a ray through the pixel centre (default; no size limits), 1 = the reference's front end: gpu cull +
                                  * rasterised visibility buffer (VulkanBaseRenderer.cpp:597-725), read like FVisibilityBufferRayCasterV2 */
} GkcuFrameParams;

/* GPUDrivenStat (assets/shaders/common/BasicTypes.slang:13-20): counters of the gpu cull, read back by Scene::UpdateNodes (Scene.cpp:412-419) */
typedef struct GkGPUDrivenStat {
    uint32_t ProcessedCount, VisibleCount, CulledCount, TriangleCount, CulledTriangleCount, pad[3];
} GkGPUDrivenStat;

typedef struct GkcuTimings {   /* names follow the reference's GPU timer scopes (PathTracingRenderer.cpp:111,126,138,160) */
    float tlas_update_ms;      /* "TLAS Update" (RayTraceBaseRenderer.cpp:209) */
    float rt_pass_ms;          /* "rt pass" */
    float reproject_pass_ms;   /* "reproject pass" */
    float compose_pass_ms;     /* "compose pass" */
    float copy_pass_ms;        /* "copy pass": pointer swaps here, ~0 */
    float frame_ms;
    float blas_build_ms;       /* last gkcu_blas_build */
    float blas_refit_ms;       /* last gkcu_blas_refit */
} GkcuTimings;

typedef struct GkcuKernelStats { /* device time of the rt-pass kernels of the last frame, CUDA events on the context stream */
    float primary_ms;          /* k_primary: primary visibility + G-buffer */
    float trace_ms;            /* sum over k_trace launches (persistent-thread BVH traversal) */
    float shade_ms;            /* sum over k_shade launches */
    float tail_ms;             /* k_tail */
    uint32_t trace_launches;
    uint32_t shade_launches;
    uint64_t trace_rays;       /* rays traversed by the k_trace launches (closest + any) */
    uint64_t tail_rays;
} GkcuKernelStats;

typedef struct GkcuCounters {  /* exact ray counts of the last frame */
    uint64_t primary_rays;
    uint64_t closest_hit_rays;
    uint64_t any_hit_rays;
    uint64_t wavefront_iterations;
    uint64_t kernel_launches;  /* kernels launched by this library during the last frame */
} GkcuCounters;

typedef struct GkcuRay { float ox, oy, oz, tmin, dx, dy, dz, tmax; } GkcuRay;
typedef struct GkcuHit { float t, u, v; uint32_t inst; uint32_t prim; } GkcuHit; /* inst = 0xFFFFFFFF on miss */

typedef struct GkcuContext GkcuContext;

/* ------------------------------------------------------------------ lifecycle */

const char* gkcu_last_error(void);

/* Replaces NextRenderer::CreateRenderer + LogicRendererBase::OnDeviceSet
 * (src/Runtime/Engine.cpp:57-90, src/Rendering/VulkanBaseRenderer.hpp:213-243). */
int gkcu_create(int device_ordinal, GkcuContext** out_ctx);
int gkcu_destroy(GkcuContext* ctx);

/* Replaces VulkanBaseRenderer::CreateRenderImages + LogicRendererBase::CreateSwapChain/DeleteSwapChain
 * (src/Rendering/VulkanBaseRenderer.cpp:429-463, PathTracingRenderer.cpp:66-102): (re)allocates the
 * storage images, zero-filled, and the wavefront path state. */
int gkcu_resize(GkcuContext* ctx, uint32_t width, uint32_t height);

/* ---------------------------------------------------------------------- scene */

/* Replaces the device uploads at the end of Scene::RebuildMeshBuffer (src/Assets/Scene.cpp:262-277). */
int gkcu_scene_upload(GkcuContext* ctx, const GkcuSceneDesc* desc);

/* Replaces RayTraceBaseRenderer::CreateBottomLevelStructures (src/Rendering/RayTraceBaseRenderer.cpp:276-340,
 * src/Vulkan/RayTracing/BottomLevelAccelerationStructure.cpp:7-59): one BLAS per model. */
int gkcu_blas_build(GkcuContext* ctx);

/* New capability (the reference never updates a BLAS): replace the positions of one model
 * (half4 per vertex, vertex_count entries, host memory) and refit its BLAS bottom-up. */
int gkcu_blas_refit(GkcuContext* ctx, uint32_t model_index, const uint16_t* simple_vertices, uint32_t vertex_count);

/* Replaces Scene::UpdateNodesGpuDriven's memcpy into the Nodes buffer + RayTraceBaseRenderer::AfterUpdateScene
 * (src/Assets/Scene.cpp:443-492, src/Rendering/RayTraceBaseRenderer.cpp:165-191): instance i <-> nodes[i]. */
int gkcu_update_nodes(GkcuContext* ctx, const GkNodeProxy* nodes, uint32_t count);

/* Replaces Scene::UpdateAllMaterials (src/Assets/Scene.cpp:393-408). */
/* Sparse variant for scenes where few nodes change per frame (the reference re-uploads the whole array every frame, Scene.cpp:443-492):
 * proxies[k] replaces node indices[k]; the node count stays what the last gkcu_update_nodes set. 208 B per changed node cross the
 * bus instead of 208 B per node; a kernel scatters them. */
int gkcu_update_nodes_sparse(GkcuContext* ctx, const uint32_t* indices, const GkNodeProxy* proxies, uint32_t count);
int gkcu_update_materials(GkcuContext* ctx, const GkMaterial* materials, uint32_t count);

/* Replaces RayTraceBaseRenderer::PreRender "TLAS Update" = TopLevelAccelerationStructure::Update, a full rebuild
 * (src/Rendering/RayTraceBaseRenderer.cpp:205-218, TopLevelAccelerationStructure.cpp:73-81). Also run implicitly
 * by gkcu_render_frame when the nodes changed since the last build. */
int gkcu_tlas_build(GkcuContext* ctx);

/* Sampled texture `id` of the bindless array (sky HDRs are ids 0-10, Engine.cpp:855): RGBA32F, row-major,
 * sampled bilinear + repeat at mip 0 (Shading.slang:194-199, src/Vulkan/Sampler.hpp:9-26). */
int gkcu_set_sky(GkcuContext* ctx, uint32_t id, const float* rgba, uint32_t width, uint32_t height);

/* GPUScene.Cubes / GPUScene.Voxels (192*48*192 entries each) or NULL,NULL for the zero probe state
 * (AmbientCube.slang:290-382). */
int gkcu_set_probes(GkcuContext* ctx, const GkAmbientCube* cubes, const GkVoxelData* voxels);

/* ---------------------------------------------------------------------- probe bake
 * GPUScene.Lights (Scene::Reload, src/Assets/Scene.cpp:92; FProcModel.cpp:110-118) and GPUScene.HDRSHs[id]
 * (GlobalTexturePool::GetHDRSphericalHarmonics, src/Assets/Texture.cpp:180-262, uploaded by Scene::UpdateHDRSH, Scene.cpp:432-441). */
int gkcu_set_lights(GkcuContext* ctx, const GkLightObject* lights, uint32_t count);
int gkcu_set_sky_sh(GkcuContext* ctx, uint32_t id, const GkSphericalHarmonics* sh);
/* Replaces FCPUAccelerationStructure::AsyncProcessFull(incremental = false) -> FCPUProbeBaker::ProcessCube(ECPT_Voxelize) over the
 * whole 192 x 48 x 192 grid and FCPUPageIndex::UpdateData (src/Assets/CPUAccelerationStructure.cpp:85-159, 313-328, 340-389,
 * 499-528): clears voxels and cubes, casts the 6 axis rays (+ 8 diagonal rays in open space) of every voxel against the current
 * TLAS, rebuilds the page index. Needs gkcu_update_nodes; builds the TLAS if it is stale. Asynchronous. */
int gkcu_probes_voxelize(GkcuContext* ctx);
/* Replaces one Bake.HwAmbientCube dispatch (assets/shaders/Bake.HwAmbientCube.comp.slang:1-20 ->
 * FGpuProbeGenerator::RenderAmbientCubeOnly, AmbientCube.slang:673-692) as recorded by RayTraceBaseRenderer::PostRender
 * (src/Rendering/RayTraceBaseRenderer.cpp:225-274): probes [first_cube, first_cube + cube_count), first_cube =
 * GPUScene.custom_data_0. The range being baked is sampled from a snapshot taken at launch (the reference races). */
int gkcu_probes_bake(GkcuContext* ctx, const GkUniformBufferObject* ubo, uint32_t first_cube, uint32_t cube_count);
/* Blocking readback of the probe state; any pointer may be NULL. pages: GK_PAGE_COUNT * GK_PAGE_COUNT entries. */
int gkcu_probes_read(GkcuContext* ctx, GkAmbientCube* cubes, GkVoxelData* voxels, GkPageIndex* pages);

/* ---------------------------------------------------------------------- frame */

/* Replaces PathTracingRenderer::Render (src/Rendering/PathTracing/PathTracingRenderer.cpp:104-218):
 * clear -> rt pass -> reproject pass -> compose pass -> copy pass (+ the object-id copy of
 * VulkanBaseRenderer::PostRender, VulkanBaseRenderer.cpp:1136-1152). Asynchronous on the context stream. */
int gkcu_render_frame(GkcuContext* ctx, const GkUniformBufferObject* ubo, const GkcuFrameParams* params);

/* Multi-GPU split of the frame: gkcu_render_frame(skip_post=1) leaves the un-normalised
 * FinalColor/FinalReflection sums in the radiance buffer (2 x W x H x float4, device memory) for an NCCL
 * reduce; gkcu_render_resolve divides by the total sample count, adds the direct term, and runs the post passes. */
int gkcu_radiance_buffer(GkcuContext* ctx, void** device_ptr, size_t* bytes);
int gkcu_render_resolve(GkcuContext* ctx, const GkUniformBufferObject* ubo, const GkcuFrameParams* params);

/* Fused reduce + resolve over peer memory (the strip variant): instead of an NCCL reduce to rank 0 followed by
 * gkcu_render_resolve there, every rank resolves its own strip of rows [rank * H / N, (rank + 1) * H / N): one kernel loads
 * that strip of every rank's radiance buffer over NVLink (peer loads), sums it in rank order, divides by the sample count,
 * adds rank 0's direct term and stores the fp16 SINGLE_DIFFUSE / SINGLE_SPECULAR strip straight into rank 0's images (peer
 * stores). Rank 0 then runs gkcu_run_post. The host orders it with two stream-ordered barriers (any tiny collective on
 * gkcu_stream): after gkcu_render_frame(skip_post = 1) and after gkcu_render_resolve_strip.
 *   gkcu_peer_export: the caller's buffer `what` as a 64-byte cudaIpcMemHandle (to ship to other processes) and as a raw
 *                     device pointer (for contexts of the same process).
 *   gkcu_peer_import: registers rank `rank`'s buffer `what`; pass the handle from another process, or handle = NULL and the
 *                     raw pointer for a context of this process. Needed: RADIANCE of every rank; DIRECT, SINGLE_DIFFUSE and
 *                     SINGLE_SPECULAR of rank 0. Imports are dropped by gkcu_resize, which also frees the buffers this rank exported:
 *                     after ANY rank of a group resized, every rank must export / import again (the barrier counters keep running).
 * If FLAGS of every rank is imported as well, gkcu_render_resolve_strip does both barriers itself on the device: each rank
 * stores a frame counter into every peer's flag block when its sums are complete / its strip has landed, and a one-warp
 * kernel polls the local block (acquire, system scope; gives up after ~2 s and reports it at gkcu_synchronize). The host then
 * issues no collective at all. */
#define GKCU_IPC_HANDLE_BYTES 64
#define GKCU_MAX_PEERS 8
enum { GKCU_PEER_RADIANCE = 0, GKCU_PEER_DIRECT = 1, GKCU_PEER_SINGLE_DIFFUSE = 2, GKCU_PEER_SINGLE_SPECULAR = 3, GKCU_PEER_FLAGS = 4,
       /* strip-parallel post (gkcu_render_resolve_post_strip): both halves of the ACCUM_* ping-pong pairs of every rank, RT_DENOISED of rank 0 */
       GKCU_PEER_ACCUM_DIFFUSE_0 = 5, GKCU_PEER_ACCUM_DIFFUSE_1 = 6, GKCU_PEER_ACCUM_SPECULAR_0 = 7, GKCU_PEER_ACCUM_SPECULAR_1 = 8,
       GKCU_PEER_ACCUM_ALBEDO_0 = 9, GKCU_PEER_ACCUM_ALBEDO_1 = 10, GKCU_PEER_DENOISED = 11, GKCU_PEER_KINDS = 12 };
int gkcu_peer_export(GkcuContext* ctx, uint32_t what, void* handle, void** device_ptr);
int gkcu_peer_import(GkcuContext* ctx, uint32_t rank, uint32_t what, const void* handle, void* device_ptr);
/* The whole tail of a multi-GPU frame on every rank, nothing left to rank 0 alone (replaces gkcu_render_resolve_strip + gkcu_run_post on rank 0):
 * device barrier -> this rank resolves its strip of rows (+ 7 halo rows either side, from every rank's radiance sums over peer memory) into its
 * OWN SINGLE_* images -> reproject pass on the strip + 5 halo rows, the owned rows also stored into every other rank's ACCUM_* (so all ranks
 * hold the full history the next frame's reprojection may look up anywhere) -> compose pass on the owned rows, stored into rank 0's
 * RT_DENOISED -> device barrier -> ping-pong swap. Halo rows are recomputed, not exchanged: identical inputs, identical bits. Needs, from
 * every rank, RADIANCE, FLAGS and the six ACCUM_* buffers, and from rank 0 DIRECT and DENOISED (gkcu_peer_import). */
int gkcu_render_resolve_post_strip(GkcuContext* ctx, const GkUniformBufferObject* ubo, const GkcuFrameParams* params);
int gkcu_render_resolve_strip(GkcuContext* ctx, const GkUniformBufferObject* ubo, const GkcuFrameParams* params);

/* Blocking device->host copy of one storage image in its reference format
 * (fp16x4 / R32_UINT / RG32F / R32F / R16_UINT, VulkanBaseRenderer.cpp:439-457). */
int gkcu_readback(GkcuContext* ctx, uint32_t rt_index, void* dst, size_t dst_bytes);
/* Presenting without stalling the renderer: enqueues the device->host copy of RT_DENOISED behind the frame submitted last and
 * returns at once. The copy runs on its own stream, so it overlaps the next frame's rt pass; the next compose pass waits for
 * it before it overwrites the image. dst must stay valid (and should be pinned) until gkcu_readback_wait has returned. */
int gkcu_readback_denoised_async(GkcuContext* ctx, void* dst, size_t dst_bytes);
int gkcu_readback_wait(GkcuContext* ctx);
/* Stencil-pass bring-up hooks used by the parity tests: overwrite one storage image from host memory (formats as in
 * gkcu_readback; PREV_* and OBJECTID_1 address the history slots), and run only reproject -> compose -> copy pass. */
int gkcu_upload_image(GkcuContext* ctx, uint32_t rt_index, const void* src, size_t src_bytes);
int gkcu_run_post(GkcuContext* ctx, const GkUniformBufferObject* ubo, const GkcuFrameParams* params);
/* Validation hooks (tests only): raw copies of the acceleration-structure pools and their bookkeeping. */
int gkcu_debug_read(GkcuContext* ctx, uint32_t what, size_t offset, void* dst, size_t bytes);
int gkcu_debug_info(GkcuContext* ctx, uint32_t* out, uint32_t capacity);
int gkcu_synchronize(GkcuContext* ctx);


/* ---------------------------------------------------------------------- visibility front end
 * gkcu_scene_upload_raster: the two buffers of Scene::RebuildMeshBuffer only the raster pass reads - the provoking index buffer
 *   (GPUScene.Indices of the draw, src/Assets/Scene.cpp:247) and GPUScene.Reorders (:240-248); ModelData.reorderOffset points into the latter.
 * gkcu_cull: "gpu cull" (assets/shaders/Task.GpuCull.comp.slang:199-227, VulkanBaseRenderer.cpp:603-647): one VkDrawIndexedIndirectCommand
 *   per NodeProxy from a frustum test and an occlusion test against last frame's RT_PREV_DEPTHBUFFER, plus the GPUDrivenStat counters.
 * gkcu_visibility_pass: "visibility pass" (Rast.VisibilityPass.{vert,frag}.slang, VulkanBaseRenderer.cpp:673-724): RT_MINIGBUFFER =
 *   ((node + 1) & 0x7FFF) << 17 | (triangle & 0x1FFFF) of the nearest front-facing triangle at every pixel centre (software rasteriser
 *   with the Vulkan rules: near clip, 8 sub-pixel bits, back-face culling, top-left rule, depth LESS). At most 32 767 proxies.
 * gkcu_render_frame runs both itself when params->primary_visibility = 1. */
int gkcu_scene_upload_raster(GkcuContext* ctx, const uint32_t* raster_indices, uint32_t index_count, const uint32_t* reorders, uint32_t reorder_count);
int gkcu_cull(GkcuContext* ctx, const GkUniformBufferObject* ubo);
int gkcu_visibility_pass(GkcuContext* ctx, const GkUniformBufferObject* ubo);
int gkcu_get_gpu_driven_stats(GkcuContext* ctx, GkGPUDrivenStat* out);
int gkcu_read_draw_commands(GkcuContext* ctx, void* dst, uint32_t count); /* count x VkDrawIndexedIndirectCommand (20 bytes) */

/* ---------------------------------------------------------------------- zero-copy present
 * The engine can create a render image whose memory is exportable as an OPAQUE_FD (RenderImage(..., external = true),
 * src/Vulkan/RenderImage.hpp:30-37; VkExternalMemoryImageCreateInfo at src/Vulkan/Image.cpp:46-58; VkExportMemoryAllocateInfo at
 * src/Vulkan/DeviceMemory.cpp:103-112 - today used for OIDN). gkcu_bind_external_image imports that memory and makes it RT_DENOISED:
 * the compose pass then writes the final image where the engine's resolve blit (VulkanBaseRenderer.cpp:1045-1069) reads it, and the
 * per-frame device->host copy of the headless path disappears. The image must be linear and tightly packed (row pitch = width * 8
 * bytes, R16G16B16A16_SFLOAT); an optimal-tiled image needs an array mapping and is not supported. handle_type:
 *   GKCU_EXT_OPAQUE_FD    a Vulkan VK_EXTERNAL_MEMORY_HANDLE_TYPE_OPAQUE_FD_BIT descriptor (cudaImportExternalMemory; CUDA owns the fd afterwards)
 *   GKCU_EXT_CUDA_VMM_FD  a cuMemExportToShareableHandle POSIX descriptor of another CUDA process / allocator (cuMemImportFromShareableHandle)
 * gkcu_bind_external_semaphore imports a VK_EXTERNAL_SEMAPHORE_HANDLE_TYPE_OPAQUE_FD semaphore that every frame signals after the compose pass
 * (fd = -1 unbinds). gkcu_resize and gkcu_destroy drop the bindings. */
enum { GKCU_EXT_OPAQUE_FD = 0, GKCU_EXT_CUDA_VMM_FD = 1 };
int gkcu_bind_external_image(GkcuContext* ctx, uint32_t rt_index, uint32_t handle_type, int fd, size_t size_bytes, size_t offset_bytes,
                             uint32_t row_pitch_bytes);
int gkcu_unbind_external_image(GkcuContext* ctx, uint32_t rt_index);
int gkcu_bind_external_semaphore(GkcuContext* ctx, uint32_t handle_type, int fd);


int gkcu_get_timings(GkcuContext* ctx, GkcuTimings* out);
int gkcu_get_counters(GkcuContext* ctx, GkcuCounters* out);
int gkcu_get_kernel_stats(GkcuContext* ctx, GkcuKernelStats* out);
/* Per-launch CUDA-event brackets around the k_trace / k_shade launches of the wavefront (GkcuKernelStats.trace_ms / shade_ms) cost about
 * 4 us per event, 0.45 ms on the 43-wave bench frame. Off (0) by default: trace_ms / shade_ms then stay 0; profiling callers (bench.py's
 * device-resident leg, the scripts under scripts/) turn them on. */
int gkcu_set_kernel_timing(GkcuContext* ctx, uint32_t enabled);
/* Compose pass (Process.DenoiseJBF.comp.slang) with the GPU's approximate exp / pow and a reciprocal instead of the pinned software
 * transcendentals and the IEEE division: RT_DENOISED then differs from the pinned build by at most 4 fp16 ulps per channel, relative RMSE
 * < 5e-4 (tests/test_gpu_parity.py::test_fast_stencil_build_stays_within_its_tolerance); the pass takes 0.08 ms instead of 0.25 ms at 1080p. Off (0) by default: the default build is bit-equal to the oracle. */
int gkcu_set_fast_stencil(GkcuContext* ctx, uint32_t enabled);
/* The reference uploads NodeProxy[] and rebuilds the TLAS every frame (Scene::UpdateNodesGpuDriven always reports a change, Scene.cpp:443-492;
 * RayTraceBaseRenderer.cpp:165-218), and so does this library by default. With this option on, gkcu_update_nodes compares the array with the
 * previous one and, when it is byte-identical, skips the upload and leaves the TLAS in place (gkcu_tlas_build then returns at once): same
 * images, 0.7 ms less per frame on RTIO.proc. */
int gkcu_set_skip_unchanged_nodes(GkcuContext* ctx, uint32_t enabled);
/* Per-frame TLAS update. The reference records a full vkCmdBuildAccelerationStructures for the TLAS every frame
 * (RayTraceBaseRenderer.cpp:205-218, TopLevelAccelerationStructure.cpp:52-80, mode BUILD). The closest hit of this library does not depend on the
 * tree, so by default (enabled = 1) gkcu_tlas_build keeps the topology of its last full build while the same instances are active and refits every
 * box from the frame's instance transforms (one single-block launch, ~0.05 ms instead of 0.7 ms for RTIO.proc's 1 936 instances); it rebuilds
 * when the active set changes, when the refitted tree's node areas have grown by 30 % over the fresh tree's, after every scene upload, and every
 * `rebuild_every` frames if that is non-zero. enabled = 0 rebuilds every frame. Images are bit-identical either way. */
int gkcu_set_tlas_refit(GkcuContext* ctx, uint32_t enabled, uint32_t rebuild_every);
/* full TLAS builds / refits since the context was created (either pointer may be NULL) */
int gkcu_tlas_stats(GkcuContext* ctx, uint64_t* builds, uint64_t* refits);
/* cudaStream_t of the context as an integer handle, for callers that time with events or enqueue collectives. */
uint64_t gkcu_stream(GkcuContext* ctx);

/* Closest-hit / any-hit queries against the current TLAS; the CUDA counterpart of
 * FCPUAccelerationStructure::RayCastInCPU (src/Assets/CPUAccelerationStructure.cpp:286-313). Host arrays. */
int gkcu_trace_rays(GkcuContext* ctx, const GkcuRay* rays, uint32_t count, GkcuHit* hits, uint32_t any_hit);

#ifdef __cplusplus
}

static_assert(sizeof(GkUniformBufferObject) == 784, "UBO layout");
static_assert(offsetof(GkUniformBufferObject, ViewportRect) == 512, "UBO layout");
static_assert(offsetof(GkUniformBufferObject, SunViewProjection) == 576, "UBO layout");
static_assert(offsetof(GkUniformBufferObject, Aperture) == 640, "UBO layout");
static_assert(offsetof(GkUniformBufferObject, TotalFrames) == 668, "UBO layout");
static_assert(offsetof(GkUniformBufferObject, NumberOfBounces) == 680, "UBO layout");
static_assert(offsetof(GkUniformBufferObject, HasSky) == 692, "UBO layout");
static_assert(offsetof(GkUniformBufferObject, TemporalFrames) == 704, "UBO layout");
static_assert(offsetof(GkUniformBufferObject, HasSun) == 708, "UBO layout");
static_assert(offsetof(GkUniformBufferObject, TAA) == 728, "UBO layout");
static_assert(offsetof(GkUniformBufferObject, ProgressiveRender) == 740, "UBO layout");
static_assert(offsetof(GkUniformBufferObject, BFSigma) == 744, "UBO layout");
static_assert(offsetof(GkUniformBufferObject, BFSize) == 756, "UBO layout");
static_assert(offsetof(GkUniformBufferObject, FastGather) == 760, "UBO layout");
static_assert(offsetof(GkUniformBufferObject, SuperResolution) == 776, "UBO layout");
static_assert(sizeof(GkNodeProxy) == 208 && offsetof(GkNodeProxy, worldTS) == 16 &&
              offsetof(GkNodeProxy, combinedPrevTS) == 80 && offsetof(GkNodeProxy, matId) == 144, "NodeProxy layout");
static_assert(sizeof(GkModelData) == 64 && offsetof(GkModelData, localAabbMin) == 16 &&
              offsetof(GkModelData, modelType) == 48 && offsetof(GkModelData, reorderOffset) == 56, "ModelData layout");
static_assert(sizeof(GkMaterial) == 64 && offsetof(GkMaterial, Fuzziness) == 28 &&
              offsetof(GkMaterial, MaterialModel) == 36 && offsetof(GkMaterial, NormalTextureScale) == 48, "Material layout");
static_assert(sizeof(GkGPUVertex) == 24, "GPUVertex layout");
static_assert(sizeof(GkAmbientCube) == 56, "AmbientCube layout");
static_assert(sizeof(GkVoxelData) == 16, "VoxelData layout");
static_assert(sizeof(GkPageIndex) == 16 && sizeof(GkSphericalHarmonics) == 112 && sizeof(GkLightObject) == 80, "probe-bake layouts");
#endif

#endif /* GKCU_H */
