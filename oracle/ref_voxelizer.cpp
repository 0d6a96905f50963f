// Compiles the reference's CPU probe voxeliser and page index as C++ from the file where they lie, to pin the oracle's restatement
// (GkoScene::VoxelizeCube / UpdatePageIndex, which the CUDA k_voxelize / k_page_index kernels are tested against):
//   src/Assets/CPUAccelerationStructure.cpp  PackBytes (25-31), DetectDistance + InsideGeometry + VoxelizeCube (68-159),
//                                            FCPUPageIndex::UpdateData + GetPage (504-553)
//   src/Assets/UniformBuffer.hpp             the grid constants (19-26)
//   assets/shaders/common/BasicTypes.slang   VoxelData, PageIndex (its C++ branch, included whole as UniformBuffer.hpp:30-48 does)
// The text is copied by oracle/Makefile into oracle/_ref/voxelizer_gen.inc unmodified. This file supplies what the extract refers to but
// which lives in the engine: the few glm types it uses (glm is not in this image), FMaterial / FetchMaterial over the oracle's material
// table, the two container structs, and TraceRay - the baker's tinybvh query (:42-66) - served by the oracle's VoxTrace (the intersector
// itself is pinned to tinybvh by liboracle_tinybvh.so). Built only when /root/reference is mounted. TEST INFRASTRUCTURE ONLY.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "gk_oracle.h"

typedef uint32_t uint;
namespace glm {
struct vec3 {
    float x, y, z;
    vec3() : x(0), y(0), z(0) {}
    template <class A, class B, class C> vec3(A a, B b, C c) : x((float)a), y((float)b), z((float)c) {}
};
inline vec3 operator*(vec3 a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
inline vec3 operator+(vec3 a, vec3 b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline vec3 operator-(vec3 a, vec3 b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline float dot(vec3 a, vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
struct vec4 { float x, y, z, w; };
struct mat4 { float m[16]; };
struct u32vec4 {
    uint32_t x, y, z, w;
    template <class A, class B, class C, class D> u32vec4(A a, B b, C c, D d) : x(static_cast<uint32_t>(a)), y(static_cast<uint32_t>(b)), z(static_cast<uint32_t>(c)), w(static_cast<uint32_t>(d)) {}
};
inline float fclamp(float v, float lo, float hi) { return std::fmin(std::fmax(v, lo), hi); }
inline int clamp(int v, int lo, int hi) { return std::min(std::max(v, lo), hi); }
namespace detail { typedef uint16_t hdata; }
}  // namespace glm

namespace Assets {
using namespace glm;
#include "uniform_buffer_consts_gen.inc"
#define float3 vec3
#define float4 vec4
#define float4x4 mat4
#define half4 glm::detail::hdata
#define public
#define bool uint32_t
#include "BasicTypes.slang"
#undef public
#undef bool
#undef half4
#undef float3
#undef float4
#undef float4x4
struct Material {  // src/Assets/Material.hpp: only what the voxeliser reads
    enum Enum : uint32_t { Lambertian = 0, Metallic = 1, Dielectric = 2, Isotropic = 3, DiffuseLight = 4, Mixture = 5 };
    uint32_t MaterialModel;
};
struct FMaterial { Material gpuMaterial_; };
}  // namespace Assets

using namespace Assets;

static GkoScene* g_scene = nullptr;
static FMaterial g_material;
FMaterial& FetchMaterial(uint matId) {  // :32-36 reads the engine's material table
    GkoSceneArrays a;
    gko_scene_arrays(g_scene, &a);
    g_material.gpuMaterial_.MaterialModel = a.materials[matId].MaterialModel;
    return g_material;
}
bool TraceRay(vec3 origin, vec3 rayDir, float dist, vec3& outNormal, uint& outMaterialId, float& outRayDist, uint& outInstanceId) {
    const float o[3] = {origin.x, origin.y, origin.z}, d[3] = {rayDir.x, rayDir.y, rayDir.z};
    float n[3];
    outInstanceId = 0;
    if (!gko_vox_trace(g_scene, o, d, dist, &outRayDist, n, &outMaterialId)) return false;
    outNormal = vec3(n[0], n[1], n[2]);
    return true;
}
struct FCPUProbeBaker { std::vector<Assets::VoxelData> voxels; };
struct FCPUPageIndex {
    std::vector<Assets::PageIndex> pageIndex;
    void UpdateData(FCPUProbeBaker& baker);
    Assets::PageIndex& GetPage(glm::vec3 worldpos);
};

#include "voxelizer_gen.inc"

extern "C" {
// FCPUProbeBaker::ProcessCube(x, y, z, ECPT_Voxelize) (:315-331) for probes [first, first + count)
void ref_voxelize(GkoScene* scene, uint32_t first, uint32_t count, void* out) {
    g_scene = scene;
    VoxelData* v = (VoxelData*)out;
    for (uint32_t i = 0; i < count; ++i) {
        const uint32_t g = first + i;
        const int y = g / (CUBE_SIZE_XY * CUBE_SIZE_XY), z = (g - y * CUBE_SIZE_XY * CUBE_SIZE_XY) / CUBE_SIZE_XY, x = g - y * CUBE_SIZE_XY * CUBE_SIZE_XY - z * CUBE_SIZE_XY;
        vec3 probePos = vec3(x, y, z) * CUBE_UNIT + CUBE_OFFSET;
        VoxelizeCube(v[i], probePos);
    }
}
void ref_page_index(const void* voxels, uint32_t n, void* pages) {
    FCPUProbeBaker baker;
    baker.voxels.assign((const VoxelData*)voxels, (const VoxelData*)voxels + n);
    FCPUPageIndex idx;
    idx.pageIndex.resize(ACGI_PAGE_COUNT * ACGI_PAGE_COUNT);  // FCPUPageIndex::Init, :499-502
    idx.UpdateData(baker);
    std::memcpy(pages, idx.pageIndex.data(), idx.pageIndex.size() * sizeof(PageIndex));
}
}
