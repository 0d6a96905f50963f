// Compiles the reference's procedural scene generators and material constructors as C++ from the files where they lie, to pin the host
// library's restatement (gknextengine_b200/host/gk_scene.cpp: CreateCornellBox / CreateBox / CreateSphere, Material*), i.e. the geometry,
// material table and quad light that BASELINE configs 1 and 2 are rendered from:
//   src/Assets/FProcModel.cpp  9-254   AddTriangle, FProcModel::CreateCornellBox, CreateBox, CreateSphere   (copied unmodified)
//   src/Assets/Material.hpp    8-81    struct Material with its constructors, struct FMaterial               (copied unmodified)
// This file supplies what those lines refer to but which lives elsewhere in the engine: the few glm types they use (glm is not in this
// image), Vertex (Vertex.hpp:10-16), LightObject (BasicTypes.slang:171-183, layout pinned by ref_layout_check), the Model constructor
// (Model.hpp:146; the MikkTSpace pass it triggers is an external library and is not reproduced) and the FProcModel declaration.
// Built only when /root/reference is mounted. TEST INFRASTRUCTURE ONLY.
#define _USE_MATH_DEFINES
#include <math.h>

#include <cmath>
#include <cstdint>
#include <cstring>
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
}  // namespace Assets

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
