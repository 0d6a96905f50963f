// Compiles the reference's dual-language header (assets/shaders/common/BasicTypes.slang, included into C++ by
// src/Assets/UniformBuffer.hpp:30-48 with the same macro shims) next to include/gkcu.h and compares sizeof/offsetof
// of every struct the C ABI mirrors. Built and run by oracle/Makefile only when /root/reference is mounted; the
// reference file is compiled where it lies and never copied. TEST INFRASTRUCTURE ONLY.
#include <cstddef>
#include <cstdint>
#include <cstdio>

#include "../include/gkcu.h"

namespace ref {
using uint = uint32_t;
struct vec3 { float x, y, z; };
struct alignas(16) vec4 { float x, y, z, w; };
struct alignas(16) mat4 { float m[16]; };
#define float3 vec3
#define float4 vec4
#define float4x4 mat4
#define half4 uint16_t
#define public
#define bool uint32_t
#include "BasicTypes.slang"
#include "BindlessTexture.slang"
#undef public
#undef bool
#undef half4
#undef float3
#undef float4
#undef float4x4
}  // namespace ref

static int fails = 0;
#define CHECK(cond) do { if (!(cond)) { std::printf("LAYOUT MISMATCH: %s\n", #cond); ++fails; } } while (0)
#define SAME_OFF(RT, GT, f) CHECK(offsetof(ref::RT, f) == offsetof(GT, f))

int main() {
    CHECK(sizeof(ref::UniformBufferObject) == sizeof(GkUniformBufferObject));
    SAME_OFF(UniformBufferObject, GkUniformBufferObject, ModelView); SAME_OFF(UniformBufferObject, GkUniformBufferObject, Projection);
    SAME_OFF(UniformBufferObject, GkUniformBufferObject, ModelViewInverse); SAME_OFF(UniformBufferObject, GkUniformBufferObject, ProjectionInverse);
    SAME_OFF(UniformBufferObject, GkUniformBufferObject, ViewProjection); SAME_OFF(UniformBufferObject, GkUniformBufferObject, PrevViewProjection);
    SAME_OFF(UniformBufferObject, GkUniformBufferObject, ViewProjectionUnJit); SAME_OFF(UniformBufferObject, GkUniformBufferObject, PrevViewProjectionUnJit);
    SAME_OFF(UniformBufferObject, GkUniformBufferObject, ViewportRect); SAME_OFF(UniformBufferObject, GkUniformBufferObject, SunDirection);
    SAME_OFF(UniformBufferObject, GkUniformBufferObject, SunColor); SAME_OFF(UniformBufferObject, GkUniformBufferObject, BackGroundColor);
    SAME_OFF(UniformBufferObject, GkUniformBufferObject, SunViewProjection); SAME_OFF(UniformBufferObject, GkUniformBufferObject, Aperture);
    SAME_OFF(UniformBufferObject, GkUniformBufferObject, FocusDistance); SAME_OFF(UniformBufferObject, GkUniformBufferObject, SkyRotation);
    SAME_OFF(UniformBufferObject, GkUniformBufferObject, HeatmapScale); SAME_OFF(UniformBufferObject, GkUniformBufferObject, PaperWhiteNit);
    SAME_OFF(UniformBufferObject, GkUniformBufferObject, SkyIntensity); SAME_OFF(UniformBufferObject, GkUniformBufferObject, SkyIdx);
    SAME_OFF(UniformBufferObject, GkUniformBufferObject, TotalFrames); SAME_OFF(UniformBufferObject, GkUniformBufferObject, MaxNumberOfBounces);
    SAME_OFF(UniformBufferObject, GkUniformBufferObject, NumberOfSamples); SAME_OFF(UniformBufferObject, GkUniformBufferObject, NumberOfBounces);
    SAME_OFF(UniformBufferObject, GkUniformBufferObject, RandomSeed); SAME_OFF(UniformBufferObject, GkUniformBufferObject, LightCount);
    SAME_OFF(UniformBufferObject, GkUniformBufferObject, HasSky); SAME_OFF(UniformBufferObject, GkUniformBufferObject, ShowHeatmap);
    SAME_OFF(UniformBufferObject, GkUniformBufferObject, UseCheckerBoard); SAME_OFF(UniformBufferObject, GkUniformBufferObject, TemporalFrames);
    SAME_OFF(UniformBufferObject, GkUniformBufferObject, HasSun); SAME_OFF(UniformBufferObject, GkUniformBufferObject, HDR);
    SAME_OFF(UniformBufferObject, GkUniformBufferObject, AdaptiveSample); SAME_OFF(UniformBufferObject, GkUniformBufferObject, AdaptiveVariance);
    SAME_OFF(UniformBufferObject, GkUniformBufferObject, AdaptiveSteps); SAME_OFF(UniformBufferObject, GkUniformBufferObject, TAA);
    SAME_OFF(UniformBufferObject, GkUniformBufferObject, SelectedId); SAME_OFF(UniformBufferObject, GkUniformBufferObject, ShowEdge);
    SAME_OFF(UniformBufferObject, GkUniformBufferObject, ProgressiveRender); SAME_OFF(UniformBufferObject, GkUniformBufferObject, BFSigma);
    SAME_OFF(UniformBufferObject, GkUniformBufferObject, BFSigmaLum); SAME_OFF(UniformBufferObject, GkUniformBufferObject, BFSigmaNormal);
    SAME_OFF(UniformBufferObject, GkUniformBufferObject, BFSize); SAME_OFF(UniformBufferObject, GkUniformBufferObject, FastGather);
    SAME_OFF(UniformBufferObject, GkUniformBufferObject, FastInterpole); SAME_OFF(UniformBufferObject, GkUniformBufferObject, DebugDraw_Lighting);
    SAME_OFF(UniformBufferObject, GkUniformBufferObject, DisableSpatialReuse); SAME_OFF(UniformBufferObject, GkUniformBufferObject, SuperResolution);

    CHECK(sizeof(ref::NodeProxy) == sizeof(GkNodeProxy));
    SAME_OFF(NodeProxy, GkNodeProxy, instanceId); SAME_OFF(NodeProxy, GkNodeProxy, modelId); SAME_OFF(NodeProxy, GkNodeProxy, visible);
    SAME_OFF(NodeProxy, GkNodeProxy, nort); SAME_OFF(NodeProxy, GkNodeProxy, worldTS); SAME_OFF(NodeProxy, GkNodeProxy, combinedPrevTS);
    SAME_OFF(NodeProxy, GkNodeProxy, matId);

    CHECK(sizeof(ref::ModelData) == sizeof(GkModelData));
    SAME_OFF(ModelData, GkModelData, indexOffset); SAME_OFF(ModelData, GkModelData, indexCount); SAME_OFF(ModelData, GkModelData, vertexOffset);
    SAME_OFF(ModelData, GkModelData, vertexCount); SAME_OFF(ModelData, GkModelData, localAabbMin); SAME_OFF(ModelData, GkModelData, localAabbMax);
    SAME_OFF(ModelData, GkModelData, modelType); SAME_OFF(ModelData, GkModelData, voxelDataIdx); SAME_OFF(ModelData, GkModelData, reorderOffset);

    CHECK(sizeof(ref::AmbientCube) == sizeof(GkAmbientCube));
    SAME_OFF(AmbientCube, GkAmbientCube, PosZ); SAME_OFF(AmbientCube, GkAmbientCube, NegX); SAME_OFF(AmbientCube, GkAmbientCube, PosZ_D);
    SAME_OFF(AmbientCube, GkAmbientCube, NegX_D); SAME_OFF(AmbientCube, GkAmbientCube, skyVisibility_pznzpyny);
    SAME_OFF(AmbientCube, GkAmbientCube, skyVisibility_pxnxs0s1);
    CHECK(sizeof(ref::VoxelData) == sizeof(GkVoxelData));
    SAME_OFF(VoxelData, GkVoxelData, matId); SAME_OFF(VoxelData, GkVoxelData, distanceToSolid_gg_z01); SAME_OFF(VoxelData, GkVoxelData, distanceToSolid_x01_y01);

    CHECK(ref::MaterialLambertian == GK_MATERIAL_LAMBERTIAN); CHECK(ref::MaterialMetallic == GK_MATERIAL_METALLIC);
    CHECK(ref::MaterialDielectric == GK_MATERIAL_DIELECTRIC); CHECK(ref::MaterialIsotropic == GK_MATERIAL_ISOTROPIC);
    CHECK(ref::MaterialDiffuseLight == GK_MATERIAL_DIFFUSE_LIGHT); CHECK(ref::MaterialMixture == GK_MATERIAL_MIXTURE);
    CHECK(ref::MAX_MATERIALS == GK_MAX_MATERIALS);

    CHECK(ref::Bindless::RT_ACCUMLATE_DIFFUSE == GK_RT_ACCUMLATE_DIFFUSE); CHECK(ref::Bindless::RT_SINGLE_DIFFUSE == GK_RT_SINGLE_DIFFUSE);
    CHECK(ref::Bindless::RT_MINIGBUFFER == GK_RT_MINIGBUFFER); CHECK(ref::Bindless::RT_OBJEDCTID_0 == GK_RT_OBJEDCTID_0);
    CHECK(ref::Bindless::RT_OBJEDCTID_1 == GK_RT_OBJEDCTID_1); CHECK(ref::Bindless::RT_MOTIONVECTOR == GK_RT_MOTIONVECTOR);
    CHECK(ref::Bindless::RT_ALBEDO == GK_RT_ALBEDO); CHECK(ref::Bindless::RT_NORMAL == GK_RT_NORMAL);
    CHECK(ref::Bindless::RT_DENOISED == GK_RT_DENOISED); CHECK(ref::Bindless::RT_PREV_DEPTHBUFFER == GK_RT_PREV_DEPTHBUFFER);
    CHECK(ref::Bindless::RT_ACCUMLATE_SPECULAR == GK_RT_ACCUMLATE_SPECULAR); CHECK(ref::Bindless::RT_SINGLE_SPECULAR == GK_RT_SINGLE_SPECULAR);
    CHECK(ref::Bindless::RT_ACCUMLATE_ALBEDO == GK_RT_ACCUMLATE_ALBEDO); CHECK(ref::Bindless::RT_MOTIONMOMENT == GK_RT_MOTIONMOMENT);
    CHECK(ref::Bindless::RT_SINGLE_PREV_DIFFUSE == GK_RT_SINGLE_PREV_DIFFUSE); CHECK(ref::Bindless::RT_SINGLE_PREV_SPECULAR == GK_RT_SINGLE_PREV_SPECULAR);
    CHECK(ref::Bindless::RT_SINGLE_PREV_ALBEDO == GK_RT_SINGLE_PREV_ALBEDO);
    std::printf(fails ? "ref_layout_check: %d mismatches\n" : "ref_layout_check: all layouts match the reference header\n", fails);
    return fails ? 1 : 0;
}
