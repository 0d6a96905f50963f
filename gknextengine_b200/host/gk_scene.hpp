// Host-side scene code above the C ABI (stays C++ like the reference).
//   * procedural fixtures: FProcModel::CreateCornellBox/CreateBox/CreateSphere (src/Assets/FProcModel.cpp:16-254)
//     and the CornellBox.proc / RTIO.proc scene lists (src/Runtime/SceneList.cpp:42-226)
//   * Scene::RebuildMeshBuffer -> GPUVertex / half4 positions / indices / ModelData (src/Assets/Scene.cpp:181-260)
//   * Scene::UpdateNodesGpuDriven + Node::TickVelocity/GetNodeProxy -> NodeProxy[] (Scene.cpp:443-492, Node.cpp:82-152)
//   * NextEngine::GetUniformBufferObject -> the 784-byte UBO (src/Runtime/Engine.cpp:642-662, 736-849)
// The outputs are exactly the arrays gkcu_scene_upload / gkcu_update_nodes / gkcu_update_materials /
// gkcu_render_frame take, and the same arrays are handed to the CPU oracle in the tests.
#pragma once
#include <array>
#include <string>
#include <vector>

#include "gk_math.hpp"

namespace gk {

struct Vertex {  // src/Assets/Vertex.hpp:10-29
    vec3 Position;
    vec3 Normal;
    vec4 Tangent;
    vec2 TexCoord;
    uint32_t MaterialIndex = 0;
};

struct Model {  // src/Assets/Model.hpp:119-170
    std::string name;
    std::vector<Vertex> vertices;
    std::vector<uint32_t> indices;
    vec3 aabb_min, aabb_max;
    uint32_t section_count = 1;
};

struct Camera {  // src/Assets/Model.hpp:12-19
    mat4 ModelView;
    float FieldOfView = 40;
    float Aperture = 0;
    float FocalDistance = 10;
};

struct EnvironmentSetting {  // src/Assets/Model.hpp:21-84
    bool HasSky = true;
    bool HasSun = false;
    int32_t SkyIdx = 0;
    float SunRotation = 0.5f;
    float SkyRotation = 0;
    float SkyIntensity = 100.0f;
    float SunIntensity = 500.0f;
    vec3 SunDirection() const;
};

struct Node {  // src/Assets/Node.cpp / Node.h
    std::string name;
    vec3 translation;
    quat rotation;
    vec3 scaling{1, 1, 1};
    uint32_t modelId = 0;
    uint32_t instanceId = 0;
    bool visible = false;
    std::array<uint32_t, 16> materialIdx{};
    mat4 transform{1.0f};
    mat4 prevTransform{1.0f};
    void RecalcTransform();
    bool TickVelocity(mat4& combinedTS);  // Node.cpp:82-106 (physics frozen: Jolt is out of scope)
};

struct UserSettings {  // src/Runtime/UserSettings.hpp + CreateUserSettings (Engine.cpp:111-163)
    int32_t NumberOfSamples = 8;
    int32_t NumberOfBounces = 5;
    int32_t MaxNumberOfBounces = 10;
    bool TAA = true;
    bool FastGather = false;
    int TemporalFrames = 16;
    bool Denoiser = false;
    float DenoiseSigma = 2.0f;
    float DenoiseSigmaLum = 3.0f;
    float DenoiseSigmaNormal = 0.005f;
    int DenoiseSize = 5;
    float PaperWhiteNit = 600.f;
    bool HDR = false;
    bool DebugDraw_Lighting = false;
    int SuperResolution = 2;
};

class Scene {
public:
    // SceneList::LoadScene (SceneList.cpp:264-291). Names: "CornellBox.proc", "RTIO.proc", plus the synthetic
    // "LegoBricks.synth[:instances]" workload of BASELINE config 4 (no reference scene file).
    bool Load(const std::string& name);

    void RebuildMeshBuffer();          // Scene.cpp:181-260
    bool UpdateNodes();                // Scene.cpp:443-492 -> node_proxies
    // Engine.cpp:736-849. `model_rotation_deg` is the benchmark orbit (ModelViewController.cpp:32-42).
    GkUniformBufferObject GetUniformBufferObject(uint32_t width, uint32_t height, uint32_t total_frames, bool progressive,
                                                 float model_rotation_deg = 0.0f);
    void ResetHistory() { has_prev_ubo_ = false; }

    // config-4 helpers: move a fraction of the instances / perturb one model's vertices (deterministic in `step`)
    void JitterInstances(uint32_t step, float fraction);
    void PerturbModel(uint32_t model, uint32_t step, std::vector<uint16_t>& out_simple_vertices);

    std::vector<Node> nodes;
    std::vector<Model> models;
    std::vector<GkMaterial> materials;
    std::vector<GkLightObject> lights;  // Scene::Lights(), filled by the procedural models (FProcModel.cpp:110-118)
    Camera camera;
    EnvironmentSetting env;
    UserSettings settings;

    // RebuildMeshBuffer outputs
    std::vector<GkGPUVertex> gpu_vertices;
    std::vector<uint16_t> simple_vertices;  // half4 per vertex
    std::vector<uint32_t> prim_indices;     // "PrimAddress" == GPUScene.Indices: model-local vertex ids, corners in provoking order
    std::vector<uint32_t> provoke_indices;  // the raster index buffer (indexes reorders through ModelData.reorderOffset), Scene.cpp:247
    std::vector<uint32_t> reorders;         // GPUScene.Reorders: absolute vertex ids, Scene.cpp:240-248
    std::vector<GkModelData> offsets;
    // UpdateNodes output
    std::vector<GkNodeProxy> node_proxies;
    // proxies that differ from the previous UpdateNodes() (a renderer may upload only these); all_proxies_changed: the array was resized
    std::vector<uint32_t> changed_proxies;
    bool all_proxies_changed = true;
    std::string last_error;  // why Load() failed

private:
    bool has_prev_ubo_ = false;
    GkUniformBufferObject prev_ubo_{};
};

GkMaterial MaterialLambertian(vec3 diffuse);
Model CreateProcModel(int kind, const float* args, std::vector<GkMaterial>& materials, std::vector<GkLightObject>& lights);
GkMaterial MakeMaterialByKind(int kind, vec3 diffuse, float a, float b);
GkMaterial MaterialMetallic(vec3 diffuse, float fuzziness);
GkMaterial MaterialMixture(vec3 diffuse, float fuzziness);
GkMaterial MaterialDielectric(float refractionIndex, float fuzziness);
GkMaterial MaterialDiffuseLight(vec3 diffuse);
// GlobalTexturePool's ProjectHdrToSh (src/Assets/Texture.cpp:180-262): 9 SH coefficients per channel of an equirect RGBA32F image.
GkSphericalHarmonics ProjectHdrToSh(const float* rgba, int width, int height);

// Procedural stand-in for the sky HDR (assets/textures/river_road_2.hdr is data that cannot travel with this
// repo): a deterministic analytic equirect RGBA32F image, rows top (+Y) to bottom.
void MakeProceduralSky(uint32_t width, uint32_t height, std::vector<float>& rgba);

// gk_gltf.cpp: FSceneLoader::LoadGLTFScene for .glb files (appends models / materials / nodes, sets camera and environment) and the
// Radiance .hdr reader behind the sky textures (stbi_loadf at src/Assets/Texture.cpp:390-432): RGBA32F, rows top to bottom
bool LoadGlbScene(Scene& s, const std::string& path, std::string& error);
bool LoadHdrImage(const std::string& path, std::vector<float>& rgba, uint32_t& width, uint32_t& height, std::string& error);

// meshopt_generateProvokingIndexBuffer restated (Scene.cpp:230), see gk_scene.cpp
size_t GenerateProvokingIndexBuffer(uint32_t* destination, uint32_t* reorder, const uint32_t* indices, size_t index_count, size_t vertex_count);

}  // namespace gk
