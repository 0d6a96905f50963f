// Compiles the reference's procedural scene generators and material constructors as C++ from the files where they lie, to pin the host
// library's restatement (gknextengine_b200/host/gk_scene.cpp: CreateCornellBox / CreateBox / CreateSphere, Material*), i.e. the geometry,
// material table and quad light that BASELINE configs 1 and 2 are rendered from:
//   src/Assets/FProcModel.cpp  9-254   AddTriangle, FProcModel::CreateCornellBox, CreateBox, CreateSphere   (copied unmodified)
//   src/Assets/Material.hpp    8-81    struct Material with its constructors, struct FMaterial               (copied unmodified)
//   src/Runtime/SceneList.cpp  27-227  CreateMaterial, AddRayTracingInOneWeekendCommonScene, RayTracingInOneWeekend, CornellBox     (copied unmodified)
//   src/Runtime/Engine.cpp     642-651 HaltonSequence (the TAA jitter of the projection, :771-777)                                    (copied unmodified)
// This file supplies what those lines refer to but which lives elsewhere in the engine: the few glm types they use (glm is not in this
// image), Vertex (Vertex.hpp:10-16), LightObject (BasicTypes.slang:171-183, layout pinned by ref_layout_check), the Model constructor
// (Model.hpp:146; the MikkTSpace pass it triggers is an external library and is not reproduced) and the FProcModel declaration.
// Built only when /root/reference is mounted. TEST INFRASTRUCTURE ONLY.
#define _USE_MATH_DEFINES
#include <math.h>

#include <array>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <memory>
#include <random>
#include <string>
#include <vector>

namespace glm {
typedef unsigned int uint;
struct vec2 {
    float x, y;
    vec2() : x(0), y(0) {}
    template <class A> explicit vec2(A v) : x((float)v), y((float)v) {}
    template <class A, class B> vec2(A a, B b) : x((float)a), y((float)b) {}
};
struct vec3 {
    float x, y, z;
    vec3() : x(0), y(0), z(0) {}
    template <class A> explicit vec3(A v) : x((float)v), y((float)v), z((float)v) {}
    template <class A, class B, class C> vec3(A a, B b, C c) : x((float)a), y((float)b), z((float)c) {}
};
inline vec3 operator-(const vec3& a, const vec3& b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline float length(const vec3& a) { return std::sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }
// rotations and the camera matrix go through glm in the engine; here only their arguments are recorded (the scene pins compare translations,
// model / instance / material ids and the material table, which is where the random stream goes)
struct quat {
    float w, x, y, z; vec3 euler; bool from_euler;
    quat(float w_, float x_, float y_, float z_) : w(w_), x(x_), y(y_), z(z_), from_euler(false) {}
    explicit quat(const vec3& e) : w(0), x(0), y(0), z(0), euler(e), from_euler(true) {}
};
struct mat4 { vec3 eye, center, up; };
inline mat4 lookAt(const vec3& eye, const vec3& center, const vec3& up) { return mat4{eye, center, up}; }
struct vec4 {
    float x, y, z, w;
    vec4() : x(0), y(0), z(0), w(0) {}
    template <class A, class B, class C, class D> vec4(A a, B b, C c, D d) : x((float)a), y((float)b), z((float)c), w((float)d) {}
    template <class D> vec4(const vec3& v, D d) : x(v.x), y(v.y), z(v.z), w((float)d) {}
};
}  // namespace glm

namespace Assets {
using namespace glm;
#include "material_gen.inc"
struct Vertex { glm::vec3 Position; glm::vec3 Normal; glm::vec4 Tangent; glm::vec2 TexCoord; glm::uint MaterialIndex; };
struct LightObject { vec4 p0, p1, p3, normal_area; uint lightMatIdx, reserved1, reserved2, reserved3; };
class Model {
public:
    Model(const std::string& name, std::vector<Vertex>&& vertices, std::vector<uint32_t>&& indices, bool needGenTSpace = true)
        : name_(name), vertices_(std::move(vertices)), indices_(std::move(indices)) {}
    std::string name_;
    std::vector<Vertex> vertices_;
    std::vector<uint32_t> indices_;
};
class FProcModel {
public:
    static Model CreateBox(const glm::vec3& p0, const glm::vec3& p1);
    static Model CreateSphere(const glm::vec3& center, float radius);
    static uint32_t CreateCornellBox(const float scale, std::vector<Model>& models, std::vector<FMaterial>& materials, std::vector<LightObject>& lights);
};
#include "procmodel_gen.inc"
struct Texture;
struct AnimationTrack {};
struct Camera { std::string name; glm::mat4 ModelView; float FieldOfView, Aperture, FocalDistance; };
struct EnvironmentSetting {  // Model.hpp:20-38, the members the scene functions set
    std::vector<Camera> cameras;
    float ControlSpeed = 5.0f;
    bool GammaCorrection = true, HasSky = true, HasSun = false;
};
}  // namespace Assets
namespace JPH { struct BodyID { uint32_t id = 0; }; enum class EMotionType { Static, Kinematic, Dynamic }; }
namespace Assets {
class Node {  // Node.h:13-62, what the scene functions call
public:
    enum class ENodeMobility { Static, Dynamic, Kinematic };
    static std::shared_ptr<Node> CreateNode(std::string name, glm::vec3 translation, glm::quat rotation, glm::vec3 scale, uint32_t modelId, uint32_t instanceId, bool replace) {
        auto n = std::make_shared<Node>();
        n->translation = translation; n->scale = scale; n->modelId = modelId; n->instanceId = instanceId;
        n->materials.fill(0u);
        return n;
    }
    void SetVisible(bool v) { visible = v; }
    void SetMaterial(const std::array<uint32_t, 16>& m) { materials = m; }
    void SetMobility(ENodeMobility) {}
    void BindPhysicsBody(JPH::BodyID) {}
    glm::vec3 translation, scale;
    uint32_t modelId = 0, instanceId = 0;
    bool visible = false;
    std::array<uint32_t, 16> materials;
};
}  // namespace Assets
struct PhysicsShim { JPH::BodyID CreateSphereBody(glm::vec3, float, JPH::EMotionType) { return JPH::BodyID{}; } };
struct NextEngine {
    static NextEngine* GetInstance() { static NextEngine e; return &e; }
    PhysicsShim* GetPhysicsEngine() { static PhysicsShim p; return &p; }
};
namespace Utilities { namespace NameHelper { inline std::string RandomName(size_t) { return "node"; } } }  // names come from std::random_device in the engine
namespace scenelist {
#include "scenelist_gen.inc"
// SceneList::LoadScene (SceneList.cpp:259-285) for the two procedural names
bool Load(const std::string& filename, Assets::EnvironmentSetting& camera, std::vector<std::shared_ptr<Assets::Node>>& nodes, std::vector<Assets::Model>& models,
          std::vector<Assets::FMaterial>& materials, std::vector<Assets::LightObject>& lights, std::vector<Assets::AnimationTrack>& tracks) {
    materials.push_back({Material::Lambertian(vec3(0.73f, 0.73f, 0.73f)), "root_default"});
    if (filename == "CornellBox.proc") { CornellBox(camera, nodes, models, materials, lights, tracks); return true; }
    if (filename == "RTIO.proc") { RayTracingInOneWeekend(camera, nodes, models, materials, lights, tracks); return true; }
    return false;
}
}  // namespace scenelist

namespace engine {
#include "halton_gen.inc"
}

using namespace Assets;
static_assert(sizeof(Material) == 64 && sizeof(LightObject) == 80, "reference layouts");

extern "C" {
// kind 0: CreateCornellBox(args[0]); 1: CreateBox(args[0..2], args[3..5]); 2: CreateSphere(args[0..2], args[3]). Same output convention as
// gkhost_proc_model: 13 words per vertex, NULL arrays to query the counts; returns the number of materials pushed.
uint32_t ref_proc_model(int kind, const float* a, void* vertices, uint32_t* vertex_count, uint32_t* indices, uint32_t* index_count, void* materials, void* light) {
    std::vector<Model> models;
    std::vector<FMaterial> mats;
    std::vector<LightObject> lights;
    if (kind == 0) FProcModel::CreateCornellBox(a[0], models, mats, lights);
    else if (kind == 1) models.push_back(FProcModel::CreateBox(vec3(a[0], a[1], a[2]), vec3(a[3], a[4], a[5])));
    else models.push_back(FProcModel::CreateSphere(vec3(a[0], a[1], a[2]), a[3]));
    const Model& m = models.back();
    *vertex_count = (uint32_t)m.vertices_.size();
    *index_count = (uint32_t)m.indices_.size();
    if (vertices) {
        uint32_t* w = static_cast<uint32_t*>(vertices);
        for (const auto& v : m.vertices_) {
            const float f[12] = {v.Position.x, v.Position.y, v.Position.z, v.Normal.x, v.Normal.y, v.Normal.z, v.Tangent.x, v.Tangent.y, v.Tangent.z, v.Tangent.w,
                                 v.TexCoord.x, v.TexCoord.y};
            std::memcpy(w, f, 48);
            w[12] = v.MaterialIndex;
            w += 13;
        }
    }
    if (indices) std::memcpy(indices, m.indices_.data(), m.indices_.size() * 4);
    if (materials)
        for (size_t i = 0; i < mats.size(); ++i) std::memcpy(static_cast<char*>(materials) + 64 * i, &mats[i].gpuMaterial_, 64);
    if (light && !lights.empty()) std::memcpy(light, &lights[0], 80);
    return (uint32_t)mats.size();
}
// The whole procedural scene: per node 24 words (translation 3 floats, scale 3 floats, modelId, instanceId, 16 material ids), the material
// table (64 bytes each), counts[0..3] = nodes, materials, models, lights; flags[0..1] = HasSky, HasSun. NULL arrays to query the counts.
int ref_scene(const char* name, uint32_t* nodes24, void* materials, uint32_t* counts, uint32_t* flags) {
    Assets::EnvironmentSetting env;
    std::vector<std::shared_ptr<Assets::Node>> nodes;
    std::vector<Model> models;
    std::vector<FMaterial> mats;
    std::vector<LightObject> lights;
    std::vector<Assets::AnimationTrack> tracks;
    if (!scenelist::Load(name, env, nodes, models, mats, lights, tracks)) return 0;
    counts[0] = (uint32_t)nodes.size(); counts[1] = (uint32_t)mats.size(); counts[2] = (uint32_t)models.size(); counts[3] = (uint32_t)lights.size();
    flags[0] = env.HasSky; flags[1] = env.HasSun;
    if (nodes24)
        for (size_t i = 0; i < nodes.size(); ++i) {
            uint32_t* w = nodes24 + 24 * i;
            const float f[6] = {nodes[i]->translation.x, nodes[i]->translation.y, nodes[i]->translation.z, nodes[i]->scale.x, nodes[i]->scale.y, nodes[i]->scale.z};
            std::memcpy(w, f, 24);
            w[6] = nodes[i]->modelId; w[7] = nodes[i]->instanceId;
            for (int k = 0; k < 16; ++k) w[8 + k] = nodes[i]->materials[k];
        }
    if (materials)
        for (size_t i = 0; i < mats.size(); ++i) std::memcpy(static_cast<char*>(materials) + 64 * i, &mats[i].gpuMaterial_, 64);
    return 1;
}
float ref_halton(int index, int base) { return engine::HaltonSequence(index, base); }
// Material::Lambertian / Metallic / Mixture / Dielectric / DiffuseLight (declaration order 0..4)
void ref_material(int kind, const float* d, float a, float b, void* out64) {
    Material m;
    const vec3 diffuse(d[0], d[1], d[2]);
    switch (kind) {
        case 0: m = Material::Lambertian(diffuse); break;
        case 1: m = Material::Metallic(diffuse, a); break;
        case 2: m = Material::Mixture(diffuse, a); break;
        case 3: m = Material::Dielectric(a, b); break;
        default: m = Material::DiffuseLight(diffuse); break;
    }
    std::memcpy(out64, &m, 64);
}
}
