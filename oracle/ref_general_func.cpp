// Compiles four more extracts of the reference's shader sources as C++ for pinning the oracle (see oracle/ref_const_func.cpp for the method):
//   Common.get_material_dataV2            assets/shaders/common/GeneralFunc.slang:6-56      
//   UnpackVertexV2, FetchMaterialId       assets/shaders/common/Const_Func.slang:356-424    (row a19)
//   mikkTSpace                            assets/shaders/common/Shading.slang:74-80         (row a14)
//   calculateWeight                       assets/shaders/Process.ReProject.comp.slang:32-47 (row a23)
// The structs they read are declared here with the reference's layouts (BasicTypes.slang:86-111, 238-254; pinned by ref_layout_check).
// Built by oracle/Makefile only when /root/reference is mounted. TEST INFRASTRUCTURE ONLY.
#include "ref_hlsl_shim.h"

namespace ref {
#include "PreProcessor.slang"
struct NodeProxy { uint instanceId, modelId, visible, nort; float4x4 worldTS, combinedPrevTS; uint matId[16]; };
struct ModelData { uint indexOffset, indexCount, vertexOffset, vertexCount; float lmin[4], lmax[4]; uint modelType, voxelDataIdx, reorderOffset, reserved2; };
struct Vertex { float3 Position, Normal; float4 Tangent; float2 TexCoord; uint MaterialIndex; float3 LocalNormal; float4 LocalTangent; };
struct GPUVertex { half4 Position_Tx, Normal_Ty, Tangent; };
static_assert(sizeof(NodeProxy) == 208 && sizeof(ModelData) == 64 && sizeof(GPUVertex) == 24, "reference layouts");
#include "general_func_gen.inc"
}  // namespace ref

using namespace ref;
extern "C" {
// out: Position 3, Normal 3, Tangent.xyz 3, TexCoord 2, LocalNormal 3, LocalTangent 4 = 18 floats; returns MaterialIndex
uint32_t ref_get_material_data(uint32_t node_plus1, uint32_t prim, const float* ro, const float* rd, const void* nodes, const void* models, const void* vertices,
                               const uint32_t* indices, float* out) {
    Vertex v = get_material_dataV2(make_uint2(node_plus1, prim), float3(ro[0], ro[1], ro[2]), float3(rd[0], rd[1], rd[2]), (NodeProxy*)nodes, (ModelData*)models,
                                   (GPUVertex*)vertices, (uint*)indices);
    const float o[18] = {v.Position.x, v.Position.y, v.Position.z, v.Normal.x, v.Normal.y, v.Normal.z, v.Tangent.x, v.Tangent.y, v.Tangent.z, v.TexCoord.x, v.TexCoord.y,
                         v.LocalNormal.x, v.LocalNormal.y, v.LocalNormal.z, v.LocalTangent.x, v.LocalTangent.y, v.LocalTangent.z, v.LocalTangent.w};
    for (int i = 0; i < 18; ++i) out[i] = o[i];
    return v.MaterialIndex;
}
uint32_t ref_fetch_material_id(const void* node, uint32_t offset) { return FetchMaterialId(*(const NodeProxy*)node, offset); }
void ref_mikktspace(const float* vnt, const float* n, const float* t, float* out) {
    float3 r = mikkTSpace(float3(vnt[0], vnt[1], vnt[2]), float3(n[0], n[1], n[2]), float4(t[0], t[1], t[2], t[3]));
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
float ref_calculate_weight(float centerDist, int sameObject, int isCenter, const float* n, const float* cn) {
    return calculateWeight(centerDist, sameObject != 0, isCenter != 0, float3(n[0], n[1], n[2]), float3(cn[0], cn[1], cn[2]));
}
}
