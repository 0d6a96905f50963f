/*
 * gk_oracle.h — CPU oracle for the gkNextEngine path-tracing frame. TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
 * library; the product (libgknext_cuda.so, gknextengine_b200/) never links, imports or calls it.
 *
 * PARITY STATUS: pinned against the reference's own shader text, compiled as C++ where it lies (oracle/Makefile target `ref`,
 * outputs only into the git-ignored oracle/_ref/, compared bit for bit by tests/test_oracle_reference.py):
 *   - the whole RT pass: Core.PathTracing + Util.BufferClear with the Shading.slang classes and interpolateAmbientCubesV2 they run
 *     (oracle/ref_pathtrace.cpp), every written storage image and both ray counts bit-identical over frames with motion
 *   - the probe bake: Bake.HwAmbientCube with FGpuProbeGenerator, EvaluateSH / SampleIBLRough (same file), cubes and ages bit-identical
 *   - the probe voxeliser and page index: line ranges of src/Assets/CPUAccelerationStructure.cpp compiled unmodified (oracle/ref_voxelizer.cpp)
 *   - the reproject and denoise passes: Process.ReProject + Process.DenoiseJBF compiled whole (oracle/ref_stencils.cpp)
 *   - the function library: Const_Func.slang, GeneralFunc.slang, UnpackVertexV2, FetchMaterialId, mikkTSpace, calculateWeight, the
 *     probe colour format and samplers (oracle/ref_const_func.cpp, oracle/ref_general_func.cpp)
 *   - struct layouts, against the reference's own dual-language header compiled as C++ (oracle/ref_layout_check.cpp)
 *   - the ray/BVH intersector, against the reference's vendored tinybvh v1.3.8 compiled from
 *     /root/reference/src/ThirdParty/tinybvh/tiny_bvh.h into oracle/_ref/ (same oracle source, -DGKO_TINYBVH)
 *   - EvaluateSH / SampleIBL (the SH sky used by the probe bake), against the reference's dual-language GLSL header compiled as
 *     C++ from /root/reference/assets/shaders/glsl_legacy/common/SampleIBL.glsl into oracle/_ref/libref_sh.so (oracle/ref_sh.cpp),
 *     and the host's ProjectHdrToSh against src/Assets/Texture.cpp:180-258 compiled unmodified (same file)
 *   - the host library's procedural models, materials and the two procedural scenes (node positions, ids, material table) against
 *     src/Assets/FProcModel.cpp, Material.hpp and src/Runtime/SceneList.cpp compiled unmodified (oracle/ref_procmodel.cpp)
 *   - fixture known-answers generated here with libstdc++, in tests/golden/.
 * "Parity unpinned" remains true only for what is hardware or driver in the reference, not source: the Vulkan driver's
 * ray/triangle rule, the raster pass behind the visibility buffer, the texture sampler and the GPU's transcendental intrinsics.
 * The reference has no tests, golden images or known-answer vectors for this path, and it cannot run here
 * (no slangc, no Vulkan).
 */
#ifndef GK_ORACLE_H
#define GK_ORACLE_H

#include "../include/gkcu.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct GkoScene GkoScene;
typedef struct GkoFrame GkoFrame;

GkoScene* gko_scene_create(const GkcuSceneDesc* desc);
void gko_scene_destroy(GkoScene* s);
void gko_update_nodes(GkoScene* s, const GkNodeProxy* nodes, uint32_t count);      /* also rebuilds the TLAS */
void gko_update_materials(GkoScene* s, const GkMaterial* materials, uint32_t count);
void gko_refit_model(GkoScene* s, uint32_t model, const uint16_t* simple_vertices, uint32_t vertex_count);
void gko_set_sky(GkoScene* s, uint32_t id, const float* rgba, uint32_t w, uint32_t h);
void gko_set_probes(GkoScene* s, const GkAmbientCube* cubes, const GkVoxelData* voxels);
const char* gko_intersector_name(void); /* "native-watertight" or "tinybvh-1.3.8" */

/* Probe grid: voxelisation + page index (src/Assets/CPUAccelerationStructure.cpp:85-159, 499-545) and one bake
 * dispatch over probes [first, first + count) (AmbientCube.slang:479-553, 673-692) with the snapshot rule of gk_shading.cuh. */
void gko_set_lights(GkoScene* s, const GkLightObject* lights, uint32_t count);
void gko_set_sky_sh(GkoScene* s, uint32_t id, const GkSphericalHarmonics* sh);
void gko_probes_voxelize(GkoScene* s, int threads);
void gko_probes_bake(GkoScene* s, const GkUniformBufferObject* ubo, uint32_t first, uint32_t count, int threads);
const GkAmbientCube* gko_probes_cubes(GkoScene* s);
const GkVoxelData* gko_probes_voxels(GkoScene* s);
const GkPageIndex* gko_probes_pages(GkoScene* s);
uint32_t gko_pack_rgb10a2(const float c[4]);
void gko_unpack_rgb10a2(uint32_t packed, float out[4]);
/* out: Position 3, Normal 3, Tangent.xyz 3, TexCoord 2, LocalNormal 3, LocalTangent 4; returns the material slot */
uint32_t gko_get_material_data(GkoScene* s, uint32_t node_plus1, uint32_t prim, const float ro[3], const float rd[3], float out[18]);
uint32_t gko_fetch_material_id(GkoScene* s, uint32_t node, uint32_t offset);
void gko_mikktspace(const float vnt[3], const float n[3], const float t[4], float out[3]);
float gko_calculate_weight(float centerDist, int sameObject, int isCenter, const float n[3], const float cn[3]);
uint32_t gko_lerp_packed_color_alt(uint32_t c0, const float c1[4], float t);
void gko_sample_ambient_cube(int which, const uint32_t cube14[14], const float n[3], float out[4]); /* 0 DI, 2 Full */
void gko_interpolate_ambient_cubes(GkoScene* s, int which, const float pos[3], const float n[3], float out[4]); /* 0 DI, 2 Full */
void gko_sample_ibl_rough(GkoScene* s, uint32_t id, const float dir[3], float rotate, float out[4]);

GkoFrame* gko_frame_create(uint32_t width, uint32_t height);
void gko_frame_destroy(GkoFrame* f);

/* One frame = clear + rt pass + reproject + compose + copies, like PathTracingRenderer::Render.
 * rows [y0, y1) bound the rt pass for timing samples (post passes then only make sense on full frames:
 * pass run_post = 0). threads <= 0 -> hardware_concurrency. */
int gko_render_frame(GkoScene* s, GkoFrame* f, const GkUniformBufferObject* ubo, const GkcuFrameParams* params,
                     int threads, uint32_t y0, uint32_t y1, int run_post);

/* Post passes alone on caller-filled images (stencil parity tests). */
int gko_reproject(GkoFrame* f, const GkcuFrameParams* params);
int gko_denoise(GkoFrame* f, const GkUniformBufferObject* ubo);
void gko_end_frame_copies(GkoFrame* f); /* ACCUM -> PREV x3, OBJECTID_0 -> OBJECTID_1 */

/* Raw storage image (reference formats), writable so tests can seed inputs. */
void* gko_image(GkoFrame* f, uint32_t rt_index, size_t* bytes);
void gko_get_counters(GkoFrame* f, GkcuCounters* out);
double gko_last_rt_seconds(GkoFrame* f);

void gko_trace_rays(GkoScene* s, const GkcuRay* rays, uint32_t count, GkcuHit* hits, uint32_t any_hit);

/* Known-answer helpers for the small pure functions (tests/test_oracle_kat.py). */
void gko_pcg4d(uint32_t v[4]);
float gko_uint_to_float(uint32_t x);
uint16_t gko_float_to_half(float f);
float gko_half_to_float(uint16_t h);
void gko_gt_tonemap(const float in[3], float out[3]);
void gko_ggx_sampling(uint32_t seed[4], float roughness, const float normal[3], float out[3]);
void gko_random_in_hemisphere1(uint32_t seed[4], float out[3]);
float gko_schlick(float cosine, float ior);
void gko_det_sincos(float x, float* s, float* c);
float gko_det_math(int fn, float a, float b); /* include/gk_detmath.h: 0 atan2(a, b), 1 acos(a), 2 exp(a), 3 log(a), 4 pow(a, b) */
void gko_st2084(const float in[3], float out[3]);
void gko_align_with_normal(const float ray[3], const float n[3], float out[3]);
float gko_random_float(uint32_t seed[4]);
void gko_random_float2(uint32_t seed[4], float out[2]);
void gko_random_in_unit_disk(uint32_t seed[4], float out[2]);
void gko_random_in_cone(uint32_t seed[4], float cos_theta, float out[3]);

/* Multi-GPU sample split: with params.skip_post = 1 gko_render_frame traces only rank params.spp_rank's samples and
 * leaves the un-normalised sums in the radiance buffer: 2 planes of W*H float4 (FinalColor, FinalReflection) followed by the
 * direct-term plane (rank 0 only, zero elsewhere). gko_resolve turns (summed) planes into SINGLE_DIFFUSE / SINGLE_SPECULAR. */
float* gko_radiance(GkoFrame* f, size_t* floats);
void gko_resolve(GkoFrame* f, const GkUniformBufferObject* ubo);

/* Services for oracle/ref_pathtrace.cpp, which compiles the reference's path-tracing shader text as C++ and needs the parts of the
 * pipeline that are hardware, not shader source: the ray query (this oracle's intersector), the bilinear texture sampler, and the
 * visibility buffer the raster pass would have written. Serial, one element per call. */
typedef struct GkoSceneArrays {
    const GkNodeProxy* nodes; const GkModelData* offsets; const GkGPUVertex* vertices; const uint32_t* indices;
    const GkMaterial* materials; uint32_t material_count, node_count;
    const GkAmbientCube* cubes; const GkVoxelData* voxels; /* NULL until gko_set_probes / gko_probes_voxelize */
    const GkLightObject* lights; const GkSphericalHarmonics* sky_sh; uint32_t light_count, sky_sh_count;
} GkoSceneArrays;
void gko_scene_arrays(GkoScene* s, GkoSceneArrays* out);
void gko_trace_one(GkoScene* s, const GkcuRay* ray, GkcuHit* hit, uint32_t any_hit);
/* the CPU baker's TraceRay (CPUAccelerationStructure.cpp:42-66): face normal through the node matrix, material of vertex 0 */
int gko_vox_trace(GkoScene* s, const float o[3], const float d[3], float dist, float* t, float normal[3], uint32_t* material);
void gko_instance_w2o(GkoScene* s, uint32_t inst, float out[12]); /* world-to-object, 3 rows x 4 columns */
void gko_sample_texture(GkoScene* s, uint32_t id, float u, float v, float out[4]);
/* packed[y*w+x] = ((node + 1) << 17) | primitive, 0 = no geometry: the RT_MINIGBUFFER texel format (Shading.slang:348-349) */
void gko_cast_visibility(GkoScene* s, const GkUniformBufferObject* ubo, uint32_t w, uint32_t h, uint32_t* packed);
/* the visibility front end (Task.GpuCull.comp.slang, Rast.VisibilityPass.*.slang): cull -> draw commands + counters, software raster -> RT_MINIGBUFFER */
void gko_set_raster_buffers(GkoScene* s, const uint32_t* raster_indices, uint32_t index_count, const uint32_t* reorders, uint32_t reorder_count);
void gko_cull(GkoScene* s, const GkUniformBufferObject* ubo, const float* prev_depth, uint32_t w, uint32_t h, uint32_t* cmds, uint32_t* stats);
void gko_visibility_pass(GkoScene* s, const GkUniformBufferObject* ubo, const uint32_t* cmds, uint32_t count, uint32_t w, uint32_t h, uint32_t* vis);
/* rows [y0, y1) only, on `threads` threads (0 = all cores) */
void gko_cast_visibility_rows(GkoScene* s, const GkUniformBufferObject* ubo, uint32_t w, uint32_t h, uint32_t y0, uint32_t y1, int threads, uint32_t* packed);

#ifdef __cplusplus
}
#endif
#endif
