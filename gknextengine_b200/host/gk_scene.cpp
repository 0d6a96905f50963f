// See gk_scene.hpp for the reference citations of each function.
#include "gk_scene.hpp"

#include <algorithm>
#include <cfloat>
#include <functional>
#include <random>

namespace gk {

// ---------------------------------------------------------------- materials (src/Assets/Material.hpp:10-38)
static GkMaterial MakeMaterial(vec4 diffuse, float fuzz, float ior, uint32_t model, float metal, float ior2 = 0.0f) {
    GkMaterial m{};
    m.Diffuse = {diffuse.x, diffuse.y, diffuse.z, diffuse.w};
    m.DiffuseTextureId = -1;
    m.MRATextureId = -1;
    m.NormalTextureId = -1;
    m.Fuzziness = fuzz;
    m.RefractionIndex = ior;
    m.MaterialModel = model;
    m.Metalness = metal;
    m.RefractionIndex2 = ior2;
    return m;
}
GkMaterial MaterialLambertian(vec3 d) { return MakeMaterial(vec4(d, 1), 1.0f, 1.0f, GK_MATERIAL_LAMBERTIAN, 0); }
GkMaterial MaterialMetallic(vec3 d, float f) { return MakeMaterial(vec4(d, 1), f, 1.45f, GK_MATERIAL_METALLIC, 1); }
GkMaterial MaterialMixture(vec3 d, float f) { return MakeMaterial(vec4(d, 1), f, 1.45f, GK_MATERIAL_MIXTURE, 0); }
GkMaterial MaterialDielectric(float ri, float f) { return MakeMaterial(vec4(1, 1, 1, 1), f, ri, GK_MATERIAL_DIELECTRIC, 0, ri); }
GkMaterial MaterialDiffuseLight(vec3 d) { return MakeMaterial(vec4(d, 1), 0.0f, 0.0f, GK_MATERIAL_DIFFUSE_LIGHT, 0); }

vec3 EnvironmentSetting::SunDirection() const {  // Model.hpp:38-41
    const float pi = 3.14159265358979323846f;
    return normalize(vec3(sinf(SunRotation * pi), 0.75f, cosf(SunRotation * pi)));
}

// ---------------------------------------------------------------- nodes
void Node::RecalcTransform() {  // Node.cpp:31-55 (no parents in the fixtures)
    transform = translate(mat4(1.0f), translation) * mat4_cast(rotation) * scale(mat4(1.0f), scaling);
}

bool Node::TickVelocity(mat4& combinedTS) {  // Node.cpp:82-106
    combinedTS = prevTransform * inverse(transform);
    prevTransform = transform;
    vec4 p = combinedTS * vec4(0, 0, 0, 1);
    bool moving = length(vec3(p.x, p.y, p.z)) > 0.001;
    if (moving) return true;
    combinedTS = mat4(1.0f);
    return false;
}

static Node CreateNode(const std::string& name, vec3 t, quat r, vec3 s, uint32_t model, uint32_t instanceId, bool replace) {
    Node n;  // Node.cpp:154-171
    n.name = name;
    n.translation = t;
    n.rotation = r;
    n.scaling = s;
    n.modelId = model;
    n.instanceId = instanceId;
    n.RecalcTransform();
    n.prevTransform = replace ? n.transform : translate(mat4(1.0f), vec3(0, -100, 0));
    return n;
}

// ---------------------------------------------------------------- procedural models
static void FinishModel(Model& m) {
    m.aabb_min = vec3(FLT_MAX);
    m.aabb_max = vec3(-FLT_MAX);
    for (auto& v : m.vertices) {
        m.aabb_min = {std::min(m.aabb_min.x, v.Position.x), std::min(m.aabb_min.y, v.Position.y), std::min(m.aabb_min.z, v.Position.z)};
        m.aabb_max = {std::max(m.aabb_max.x, v.Position.x), std::max(m.aabb_max.y, v.Position.y), std::max(m.aabb_max.z, v.Position.z)};
    }
}
static Vertex V(vec3 p, vec3 n, float u, float v, uint32_t mat) {
    Vertex r;
    r.Position = p;
    r.Normal = n;
    r.Tangent = vec4(1, 0, 0, 0);
    r.TexCoord = {u, v};
    r.MaterialIndex = mat;
    return r;
}
static void Quad(Model& m, const Vertex& a, const Vertex& b, const Vertex& c, const Vertex& d, const uint32_t (&tri)[6]) {
    uint32_t i = (uint32_t)m.vertices.size();
    m.vertices.push_back(a); m.vertices.push_back(b); m.vertices.push_back(c); m.vertices.push_back(d);
    for (uint32_t k : tri) m.indices.push_back(i + k);
}

// FProcModel::CreateCornellBox (FProcModel.cpp:16-128): pushes 4 materials, returns the model.
static Model CreateCornellBox(float s, std::vector<GkMaterial>& materials, std::vector<GkLightObject>& lights) {
    const uint32_t prevMatId = (uint32_t)materials.size();
    materials.push_back(MaterialLambertian(vec3(0.65f, 0.05f, 0.05f)));  // red
    materials.push_back(MaterialLambertian(vec3(0.12f, 0.45f, 0.15f)));  // green
    materials.push_back(MaterialLambertian(vec3(0.73f, 0.73f, 0.73f)));  // white
    materials.push_back(MaterialDiffuseLight(vec3(2000.0f)));            // light
    Model m;
    m.name = "cornell_box";
    const vec3 offset(s * 0.5f, 0, -s * 0.5f);
    const vec3 l0(0, 0, 0), l1(0, 0, -s), l2(0, s, -s), l3(0, s, 0);
    const vec3 r0(s, 0, 0), r1(s, 0, -s), r2(s, s, -s), r3(s, s, 0);
    const uint32_t fwd[6] = {0, 1, 2, 0, 2, 3}, rev[6] = {2, 1, 0, 3, 2, 0};
    vec3 n;
    n = vec3(1, 0, 0);   // left panel, slot 1
    Quad(m, V(l0 - offset, n, 0, 1, 1), V(l1 - offset, n, 1, 1, 1), V(l2 - offset, n, 1, 0, 1), V(l3 - offset, n, 0, 0, 1), fwd);
    n = vec3(-1, 0, 0);  // right panel, slot 0
    Quad(m, V(r0 - offset, n, 0, 1, 0), V(r1 - offset, n, 1, 1, 0), V(r2 - offset, n, 1, 0, 0), V(r3 - offset, n, 0, 0, 0), rev);
    n = vec3(0, 0, 1);   // back
    Quad(m, V(l1 - offset, n, 0, 1, 2), V(r1 - offset, n, 1, 1, 2), V(r2 - offset, n, 1, 0, 2), V(l2 - offset, n, 0, 0, 2), fwd);
    n = vec3(0, 1, 0);   // bottom
    Quad(m, V(l0 - offset, n, 0, 1, 2), V(r0 - offset, n, 1, 1, 2), V(r1 - offset, n, 1, 0, 2), V(l1 - offset, n, 0, 0, 2), fwd);
    n = vec3(0, -1, 0);  // top
    Quad(m, V(l2 - offset, n, 0, 1, 2), V(r2 - offset, n, 1, 1, 2), V(r3 - offset, n, 1, 0, 2), V(l3 - offset, n, 0, 0, 2), fwd);
    {  // light
        const float x0 = s * (163.0f / 555.0f), x1 = s * (393.0f / 555.0f);
        const float z0 = s * (-555.0f + 432.0f) / 555.0f, z1 = s * (-555.0f + 202.0f) / 555.0f;
        const float y1 = s * 0.999f;
        Quad(m, V(vec3(x0, y1, z1) - offset, n, 0, 1, 3), V(vec3(x1, y1, z1) - offset, n, 1, 1, 3),
             V(vec3(x1, y1, z0) - offset, n, 1, 0, 3), V(vec3(x0, y1, z0) - offset, n, 0, 0, 3), fwd);
        GkLightObject light{};  // FProcModel.cpp:110-118
        const vec3 p0 = vec3(x0, y1, z1) - offset, p1 = vec3(x0, y1, z0) - offset, p3 = vec3(x1, y1, z1) - offset;
        light.p0 = {p0.x, p0.y, p0.z, 1};
        light.p1 = {p1.x, p1.y, p1.z, 1};
        light.p3 = {p3.x, p3.y, p3.z, 1};
        light.normal_area = {0, -1, 0, (x1 - x0) * (z0 - z1)};
        light.lightMatIdx = prevMatId + 3;
        lights.push_back(light);
    }
    FinishModel(m);
    return m;
}

// FProcModel::CreateBox (FProcModel.cpp:130-176)
static Model CreateBox(vec3 p0, vec3 p1) {
    Model m;
    m.name = "pbox";
    const uint32_t fwd[6] = {0, 1, 2, 0, 2, 3};
    vec3 n;
    n = vec3(-1, 0, 0);
    Quad(m, V(vec3(p0.x, p0.y, p0.z), n, 0, 0, 0), V(vec3(p0.x, p0.y, p1.z), n, 0, 0, 0), V(vec3(p0.x, p1.y, p1.z), n, 0, 0, 0), V(vec3(p0.x, p1.y, p0.z), n, 0, 0, 0), fwd);
    n = vec3(1, 0, 0);
    Quad(m, V(vec3(p1.x, p0.y, p1.z), n, 0, 0, 0), V(vec3(p1.x, p0.y, p0.z), n, 0, 0, 0), V(vec3(p1.x, p1.y, p0.z), n, 0, 0, 0), V(vec3(p1.x, p1.y, p1.z), n, 0, 0, 0), fwd);
    n = vec3(0, 0, -1);
    Quad(m, V(vec3(p1.x, p0.y, p0.z), n, 0, 0, 0), V(vec3(p0.x, p0.y, p0.z), n, 0, 0, 0), V(vec3(p0.x, p1.y, p0.z), n, 0, 0, 0), V(vec3(p1.x, p1.y, p0.z), n, 0, 0, 0), fwd);
    n = vec3(0, 0, 1);
    Quad(m, V(vec3(p0.x, p0.y, p1.z), n, 0, 0, 0), V(vec3(p1.x, p0.y, p1.z), n, 0, 0, 0), V(vec3(p1.x, p1.y, p1.z), n, 0, 0, 0), V(vec3(p0.x, p1.y, p1.z), n, 0, 0, 0), fwd);
    n = vec3(0, -1, 0);
    Quad(m, V(vec3(p0.x, p0.y, p0.z), n, 0, 0, 0), V(vec3(p1.x, p0.y, p0.z), n, 0, 0, 0), V(vec3(p1.x, p0.y, p1.z), n, 0, 0, 0), V(vec3(p0.x, p0.y, p1.z), n, 0, 0, 0), fwd);
    n = vec3(0, 1, 0);
    Quad(m, V(vec3(p1.x, p1.y, p0.z), n, 0, 0, 0), V(vec3(p0.x, p1.y, p0.z), n, 0, 0, 0), V(vec3(p0.x, p1.y, p1.z), n, 0, 0, 0), V(vec3(p1.x, p1.y, p1.z), n, 0, 0, 0), fwd);
    FinishModel(m);
    return m;
}

// FProcModel::CreateSphere (FProcModel.cpp:178-254): 32 slices x 16 stacks, 561 vertices, 1024 triangles.
static Model CreateSphere(vec3 center, float radius) {
    const int slices = 32, stacks = 16;
    const float pi = float(3.14159265358979323846);
    Model m;
    m.name = "sphere";
    const float j0Delta = pi / static_cast<float>(stacks);
    float j0 = 0.f;
    const float i0Delta = (pi + pi) / static_cast<float>(slices);
    for (int j = 0; j <= stacks; ++j) {
        const float v = radius * -std::sin(j0);
        const float z = radius * std::cos(j0);
        const float n0 = -std::sin(j0);
        const float n1 = std::cos(j0);
        float i0 = 0;
        for (int i = 0; i <= slices; ++i) {
            Vertex vert;
            vert.Position = vec3(center.x + v * std::sin(i0), center.y + z, center.z + v * std::cos(i0));
            vert.Normal = vec3(n0 * std::sin(i0), n1, n0 * std::cos(i0));
            vert.Tangent = vec4(1, 0, 0, 0);
            vert.TexCoord = {static_cast<float>(i) / slices, static_cast<float>(j) / stacks};
            vert.MaterialIndex = 0;
            m.vertices.push_back(vert);
            i0 += i0Delta;
        }
        j0 += j0Delta;
    }
    const int slices1 = slices + 1;
    int r0 = 0, r1 = slices1;
    for (int j = 0; j < stacks; ++j) {
        for (int i = 0; i < slices; ++i) {
            const uint32_t a = i, b = i + 1;
            m.indices.push_back(r0 + a); m.indices.push_back(r1 + a); m.indices.push_back(r1 + b);
            m.indices.push_back(r0 + a); m.indices.push_back(r1 + b); m.indices.push_back(r0 + b);
        }
        r0 += slices1;
        r1 += slices1;
    }
    FinishModel(m);
    return m;
}

// A lego-style brick: w x d studs on a box, 6-sided studs (12 + 16*w*d triangles). Config-4 workload only.
static Model CreateBrick(int w, int d) {
    const float unit = 0.2f, height = 0.24f, stud_r = 0.06f, stud_h = 0.04f;
    Model m = CreateBox(vec3(0, 0, 0), vec3(w * unit, height, d * unit));
    m.name = "brick";
    const float pi = 3.14159265358979323846f;
    for (int sx = 0; sx < w; ++sx)
        for (int sz = 0; sz < d; ++sz) {
            vec3 c((sx + 0.5f) * unit, height, (sz + 0.5f) * unit);
            uint32_t base = (uint32_t)m.vertices.size();
            for (int k = 0; k < 6; ++k) {  // side ring: bottom/top pairs with radial normals
                float a = 2 * pi * k / 6;
                vec3 n(std::cos(a), 0, std::sin(a));
                m.vertices.push_back(V(c + n * stud_r, n, 0, 0, 0));
                m.vertices.push_back(V(c + n * stud_r + vec3(0, stud_h, 0), n, 0, 0, 0));
            }
            for (int k = 0; k < 6; ++k) {
                uint32_t a0 = base + 2 * k, a1 = a0 + 1, b0 = base + 2 * ((k + 1) % 6), b1 = b0 + 1;
                m.indices.insert(m.indices.end(), {a0, a1, b1, a0, b1, b0});
            }
            uint32_t cap = (uint32_t)m.vertices.size();
            for (int k = 0; k < 6; ++k) {
                float a = 2 * pi * k / 6;
                m.vertices.push_back(V(c + vec3(std::cos(a) * stud_r, stud_h, std::sin(a) * stud_r), vec3(0, 1, 0), 0, 0, 0));
            }
            for (int k = 1; k < 5; ++k) m.indices.insert(m.indices.end(), {cap, cap + (uint32_t)k + 1, cap + (uint32_t)k});
        }
    FinishModel(m);
    return m;
}

// ---------------------------------------------------------------- scene lists
static void AddNode(Scene& s, Node n, std::initializer_list<uint32_t> mats) {
    n.visible = true;
    int i = 0;
    for (uint32_t m : mats) n.materialIdx[i++] = m;
    s.nodes.push_back(n);
}

static void LoadCornellBox(Scene& s) {  // SceneList.cpp:180-226
    uint32_t prevMatId = (uint32_t)s.materials.size();
    s.camera.ModelView = look_at_rh(vec3(0, 2.78f, 10.78f), vec3(0, 2.78f, 0), vec3(0, 1, 0));
    s.camera.FieldOfView = 40;
    s.camera.Aperture = 0;
    s.camera.FocalDistance = 10;
    s.env.HasSky = false;
    s.env.HasSun = false;
    s.models.push_back(CreateCornellBox(5.55f, s.materials, s.lights));
    uint32_t cboxModel = (uint32_t)s.models.size() - 1;
    AddNode(s, CreateNode("cbox", vec3(0, 0, 0), quat{1, 0, 0, 0}, vec3(1, 1, 1), cboxModel, (uint32_t)s.nodes.size(), false),
            {prevMatId + 0, prevMatId + 1, prevMatId + 2, prevMatId + 3});
    vec3 spherePos(1.30f, 1.01f + 2.00f * 0.0f, 0.80f);
    vec3 boxPos(-1.30f, 0, -0.80f);
    s.materials.push_back(MaterialLambertian(vec3(0.73f, 0.73f, 0.73f)));     // cbox_white
    s.materials.push_back(MaterialMixture(vec3(0.73f, 0.73f, 0.73f), 0.01f)); // cball_white
    s.models.push_back(CreateBox(vec3(-0.80f, 0, -0.80f), vec3(0.80f, 1.60f, 0.80f)));
    s.models.push_back(CreateSphere(vec3(0, 0, 0), 1.0f));
    AddNode(s, CreateNode("Sphere1", spherePos, quat_from_euler(vec3(0, 0.5f, 0)), vec3(1, 1, 1), cboxModel + 2, (uint32_t)s.nodes.size(), false),
            {prevMatId + 5});
    AddNode(s, CreateNode("Box", boxPos, quat_from_euler(vec3(0, 0.25f, 0)), vec3(1, 2, 1), cboxModel + 1, (uint32_t)s.nodes.size(), false),
            {prevMatId + 4});
}

static void LoadRTIOW(Scene& s) {  // SceneList.cpp:42-178
    uint32_t prevMatId = (uint32_t)s.materials.size();
    s.camera.ModelView = look_at_rh(vec3(13, 2, 3), vec3(0, 0, 0), vec3(0, 1, 0));
    s.camera.FieldOfView = 20;
    s.camera.Aperture = 0;
    s.camera.FocalDistance = 10;
    s.env.HasSky = true;
    s.env.HasSun = false;

    std::mt19937 engine(42);
    std::function<float()> random = std::bind(std::uniform_real_distribution<float>(), engine);

    s.materials.push_back(MaterialLambertian(vec3(0.4f, 0.4f, 0.4f)));
    s.models.push_back(CreateBox(vec3(-1000, -0.5f, -1000), vec3(1000, 0, 1000)));
    AddNode(s, CreateNode("floor", vec3(0, 0, 0), quat{1, 0, 0, 0}, vec3(1, 1, 1), 0, (uint32_t)s.nodes.size(), false), {prevMatId + 0});

    s.models.push_back(CreateSphere(vec3(0, 0, 0), 0.2f));
    uint32_t meshIdx = (uint32_t)s.models.size() - 1;
    for (int i = -22; i < 22; ++i) {
        for (int j = -22; j < 22; ++j) {
            const float chooseMat = random();
            const float centerY = static_cast<float>(j) + 0.9f * random();
            const float centerX = static_cast<float>(i) + 0.9f * random();
            const vec3 center(centerX, 0.2f + 2.0f * chooseMat, centerY);
            if (length(center - vec3(4, 0.2f, 0)) > 0.9f) {
                uint32_t matId;
                if (chooseMat < 0.7f) {
                    const float b = random() * random();
                    const float g = random() * random();
                    const float r = random() * random();
                    s.materials.push_back(MaterialLambertian(vec3(r, g, b)));
                } else if (chooseMat < 0.9f) {
                    const float fuzziness = 0.5f * random();
                    const float b = 0.5f * (1 + random());
                    const float g = 0.5f * (1 + random());
                    const float r = 0.5f * (1 + random());
                    s.materials.push_back(MaterialMetallic(vec3(r, g, b), fuzziness));
                } else {
                    const float fuzziness = 0.5f * random();
                    s.materials.push_back(MaterialDielectric(1.5f, fuzziness));
                }
                matId = (uint32_t)s.materials.size() - 1;
                AddNode(s, CreateNode("s", center, quat{1, 0, 0, 0}, vec3(1, 1, 1), meshIdx, (uint32_t)s.nodes.size(), false), {matId});
            }
        }
    }
    s.materials.push_back(MaterialDielectric(1.5f, 0.0f));
    uint32_t matIdx0 = (uint32_t)s.materials.size() - 1;
    s.materials.push_back(MaterialLambertian(vec3(0.4f, 0.2f, 0.1f)));
    uint32_t matIdx1 = (uint32_t)s.materials.size() - 1;
    s.materials.push_back(MaterialMetallic(vec3(0.7f, 0.6f, 0.5f), 0.1f));
    uint32_t matIdx2 = (uint32_t)s.materials.size() - 1;
    s.models.push_back(CreateSphere(vec3(0, 0, 0), 1.0f));
    uint32_t modelIdx = (uint32_t)s.models.size() - 1;
    AddNode(s, CreateNode("b0", vec3(0, 1, 0), quat{1, 0, 0, 0}, vec3(1, 1, 1), modelIdx, (uint32_t)s.nodes.size(), false), {matIdx0});
    AddNode(s, CreateNode("b1", vec3(-4, 1, 0), quat{1, 0, 0, 0}, vec3(1, 1, 1), modelIdx, (uint32_t)s.nodes.size(), false), {matIdx1});
    AddNode(s, CreateNode("b2", vec3(4, 1, 0), quat{1, 0, 0, 0}, vec3(1, 1, 1), modelIdx, (uint32_t)s.nodes.size(), false), {matIdx2});
}

// Config 4: 16 brick models, `count` instances on an integer lattice (mt19937(1234)), sky + sun.
static void LoadLegoSynth(Scene& s, uint32_t count) {
    s.camera.FieldOfView = 40;
    s.camera.Aperture = 0;
    s.camera.FocalDistance = 10;
    s.env.HasSky = true;
    s.env.HasSun = true;
    s.materials.push_back(MaterialLambertian(vec3(0.5f, 0.5f, 0.5f)));
    uint32_t floorMat = (uint32_t)s.materials.size() - 1;
    const vec3 palette[8] = {vec3(0.8f, 0.1f, 0.1f), vec3(0.1f, 0.3f, 0.8f), vec3(0.9f, 0.8f, 0.1f), vec3(0.1f, 0.6f, 0.2f),
                             vec3(0.9f, 0.9f, 0.9f), vec3(0.1f, 0.1f, 0.1f), vec3(0.9f, 0.5f, 0.1f), vec3(0.5f, 0.2f, 0.6f)};
    uint32_t firstColor = (uint32_t)s.materials.size();
    for (int k = 0; k < 8; ++k) s.materials.push_back(MaterialMixture(palette[k], 0.2f));
    uint32_t side = (uint32_t)std::ceil(std::sqrt((double)count / 4.0));  // ~4 layers high
    const float unit = 0.2f, height = 0.24f;
    float extent = side * 4 * unit;
    s.models.push_back(CreateBox(vec3(-extent, -0.5f, -extent), vec3(extent, 0, extent)));
    AddNode(s, CreateNode("floor", vec3(0, 0, 0), quat{1, 0, 0, 0}, vec3(1, 1, 1), 0, 0, true), {floorMat});
    uint32_t firstBrick = (uint32_t)s.models.size();
    for (int w = 1; w <= 4; ++w)
        for (int d = 1; d <= 4; ++d) s.models.push_back(CreateBrick(w, d));
    std::mt19937 rng(1234);
    for (uint32_t i = 0; i < count; ++i) {
        uint32_t cell = i / 4, layer = i % 4;
        int gx = (int)(cell % side) - (int)side / 2, gz = (int)(cell / side) - (int)side / 2;
        uint32_t model = firstBrick + rng() % 16;
        uint32_t mat = firstColor + rng() % 8;
        uint32_t rot = rng() % 4;
        vec3 pos(gx * 4 * unit, layer * height, gz * 4 * unit);
        AddNode(s, CreateNode("brick", pos, quat_from_euler(vec3(0, rot * 1.57079632679f, 0)), vec3(1, 1, 1), model,
                              (uint32_t)s.nodes.size(), true), {mat});
    }
    float r = extent * 0.55f;
    s.camera.ModelView = look_at_rh(vec3(r, r * 0.6f, r), vec3(0, 0, 0), vec3(0, 1, 0));
}

bool Scene::Load(const std::string& name) {
    nodes.clear(); models.clear(); materials.clear();
    env = EnvironmentSetting();
    has_prev_ubo_ = false;
    materials.push_back(MaterialLambertian(vec3(0.73f, 0.73f, 0.73f)));  // root_default (SceneList.cpp:270)
    if (name == "CornellBox.proc") LoadCornellBox(*this);
    else if (name == "RTIO.proc") LoadRTIOW(*this);
    else if (name.rfind("LegoBricks.synth", 0) == 0) {
        uint32_t count = 100000;
        auto pos = name.find(':');
        if (pos != std::string::npos) count = (uint32_t)std::stoul(name.substr(pos + 1));
        LoadLegoSynth(*this, count);
    } else if (name.size() > 4 && name.compare(name.size() - 4, 4, ".glb") == 0) {
        // SceneList::LoadScene -> FSceneLoader::LoadGLTFScene (SceneList.cpp:264-291): a path, or a file name looked up in
        // $GKCU_ASSET_DIR and in the reference's assets/models when that tree is mounted
        std::string err, path = name;
        auto exists = [](const std::string& p) { FILE* f = std::fopen(p.c_str(), "rb"); if (f) std::fclose(f); return f != nullptr; };
        if (!exists(path) && std::getenv("GKCU_ASSET_DIR")) path = std::string(std::getenv("GKCU_ASSET_DIR")) + "/" + name;
        if (!exists(path)) path = "/root/reference/assets/models/" + name;
        if (!LoadGlbScene(*this, path, err)) { last_error = err; return false; }
    } else return false;
    RebuildMeshBuffer();
    UpdateNodes();
    return true;
}

// ---------------------------------------------------------------- mesh buffers (Scene.cpp:181-260, Vertex.hpp:80-99)
static GkGPUVertex MakeVertex(const Vertex& v) {
    GkGPUVertex g;
    g.posx = to_float16(v.Position.x); g.posy = to_float16(v.Position.y); g.posz = to_float16(v.Position.z);
    g.texcoordx = to_float16(v.TexCoord.x);
    g.normalx = to_float16(v.Normal.x); g.normaly = to_float16(v.Normal.y); g.normalz = to_float16(v.Normal.z);
    g.texcoordy = to_float16(v.TexCoord.y);
    g.tangentx = to_float16(v.Tangent.x); g.tangenty = to_float16(v.Tangent.y); g.tangentz = to_float16(v.Tangent.z);
    g.tangentw = (uint16_t)(((v.Tangent.w > 0 ? 2 : 0) << 8) | (uint16_t)v.MaterialIndex);
    return g;
}

// meshopt_generateProvokingIndexBuffer (meshoptimizer >= 0.22, src/indexgenerator.cpp; a vcpkg dependency of the reference that is not
// in the mount, called at Scene.cpp:230), restated from the published algorithm: every triangle is rotated so that its first
// ("provoking") vertex is one no earlier triangle provoked - among the unseen corners the one of least remaining valence - and gets a
// fresh slot in `reorder`; the other two corners are remapped to the slot of the vertex they reference (cloned when that vertex never
// provoked). destination[] indexes reorder[], reorder[] holds original vertex ids. Returns the number of reorder entries.
size_t GenerateProvokingIndexBuffer(uint32_t* destination, uint32_t* reorder, const uint32_t* indices, size_t index_count, size_t vertex_count) {
    std::vector<uint32_t> remap(vertex_count, ~0u);
    std::vector<uint8_t> valence(vertex_count, 0);  // 8-bit counters like the original: an overflow only perturbs the heuristic, and identically
    for (size_t i = 0; i < index_count; ++i) valence[indices[i]]++;
    uint32_t reorder_offset = 0;
    for (size_t i = 0; i < index_count; i += 3) {
        uint32_t a = indices[i + 0], b = indices[i + 1], c = indices[i + 2];
        const uint32_t va = remap[a] == ~0u ? valence[a] : ~0u;
        const uint32_t vb = remap[b] == ~0u ? valence[b] : ~0u;
        const uint32_t vc = remap[c] == ~0u ? valence[c] : ~0u;
        if (vb != ~0u && vb <= va && vb <= vc) {  // abc -> bca
            const uint32_t t = a;
            a = b, b = c, c = t;
        } else if (vc != ~0u && vc <= va && vc <= vb) {  // abc -> cab
            const uint32_t t = c;
            c = b, b = a, a = t;
        }
        const uint32_t newidx = reorder_offset;
        if (remap[a] == ~0u) remap[a] = newidx;
        reorder[reorder_offset++] = a;
        destination[i + 0] = newidx;
        destination[i + 1] = b;
        destination[i + 2] = c;
        valence[a]--; valence[b]--; valence[c]--;
    }
    int step = 1;
    for (size_t i = 1; i < index_count; i += step, step ^= 3) {  // the two non-provoking corners of every triangle
        const uint32_t index = destination[i];
        if (remap[index] == ~0u) {
            remap[index] = reorder_offset;
            reorder[reorder_offset++] = index;
        }
        destination[i] = remap[index];
    }
    return reorder_offset;
}

void Scene::RebuildMeshBuffer() {
    gpu_vertices.clear(); simple_vertices.clear(); prim_indices.clear(); offsets.clear(); provoke_indices.clear(); reorders.clear();
    for (auto& model : models) {
        const uint32_t vertexOffset = (uint32_t)gpu_vertices.size();
        for (auto& v : model.vertices) {
            gpu_vertices.push_back(MakeVertex(v));
            simple_vertices.push_back(to_float16(v.Position.x));
            simple_vertices.push_back(to_float16(v.Position.y));
            simple_vertices.push_back(to_float16(v.Position.z));
            simple_vertices.push_back(to_float16(v.Position.x));
        }
        // <= 65535-triangle sections, 10 ModelData slots per model (Scene.cpp:204-256). Every section goes through the provoking-vertex
        // index generator (Scene.cpp:224-250): `provoke_indices` (GPUScene's raster index buffer) indexes `reorders` (absolute vertex
        // ids), and `prim_indices` = reorder[provoke[i]] are the model-local vertex ids in the ROTATED corner order - the buffer the
        // shaders call Indices and the BLAS is built from.
        const uint32_t maxIndicesPerSlice = 65535 * 3;
        uint32_t total = (uint32_t)model.indices.size();
        uint32_t slices = (total + maxIndicesPerSlice - 1) / maxIndicesPerSlice;
        uint32_t process = std::min<uint32_t>(slices, 10);
        const uint32_t modelIndexOffset = (uint32_t)prim_indices.size(), modelReorderOffset = (uint32_t)reorders.size();
        auto md = [&](uint32_t indexOffset, uint32_t count, uint32_t reorderOffset) {
            GkModelData d{};
            d.indexOffset = indexOffset; d.indexCount = count; d.vertexOffset = vertexOffset; d.vertexCount = (uint32_t)model.vertices.size();
            d.localAabbMin = {model.aabb_min.x, model.aabb_min.y, model.aabb_min.z, 1};
            d.localAabbMax = {model.aabb_max.x, model.aabb_max.y, model.aabb_max.z, 1};
            d.reorderOffset = reorderOffset;
            return d;
        };
        for (uint32_t sl = 0; sl < process; ++sl) {
            const uint32_t begin = sl * maxIndicesPerSlice, end = std::min(begin + maxIndicesPerSlice, total), n = end - begin;
            offsets.push_back(md((uint32_t)prim_indices.size(), n, (uint32_t)reorders.size()));
            std::vector<uint32_t> provoke(n), reorder(model.vertices.size() + n / 3);
            if (n > 0) reorder.resize(GenerateProvokingIndexBuffer(provoke.data(), reorder.data(), model.indices.data() + begin, n, model.vertices.size()));
            else reorder.clear();
            for (uint32_t i = 0; i < n; ++i) prim_indices.push_back(reorder[provoke[i]]);
            provoke_indices.insert(provoke_indices.end(), provoke.begin(), provoke.end());
            for (uint32_t r : reorder) reorders.push_back(r + vertexOffset);  // "reorder is absolute vertex index"
        }
        for (uint32_t sl = process; sl < 10; ++sl) offsets.push_back(md(modelIndexOffset, 0, modelReorderOffset));
        model.section_count = std::max<uint32_t>(process, 1);
    }
}

bool Scene::UpdateNodes() {  // Scene.cpp:443-492 + Node::GetNodeProxy (Node.cpp:136-152)
    std::vector<GkNodeProxy> previous;
    previous.swap(node_proxies);
    node_proxies.reserve(previous.size());
    for (auto& node : nodes) {
        if (node.modelId == 0xFFFFFFFFu) continue;  // Node::IsDrawable(): invisible nodes keep their proxy (visible = 0 masks the TLAS instance)
        mat4 combined;
        node.TickVelocity(combined);
        const Model& model = models[node.modelId];
        for (uint32_t section = 0; section < model.section_count; ++section) {
            GkNodeProxy p{};
            p.instanceId = node.instanceId;
            p.modelId = node.modelId * 10 + section;
            p.visible = node.visible ? 1 : 0;
            p.nort = section == 0 ? 0 : 1;
            p.worldTS = node.transform.pod();
            p.combinedPrevTS = combined.pod();
            for (int i = 0; i < 16; ++i) p.matId[i] = node.materialIdx[i];
            node_proxies.push_back(p);
        }
    }
    changed_proxies.clear();
    all_proxies_changed = previous.size() != node_proxies.size();
    if (!all_proxies_changed)
        for (size_t i = 0; i < node_proxies.size(); ++i)
            if (std::memcmp(&previous[i], &node_proxies[i], sizeof(GkNodeProxy)) != 0) changed_proxies.push_back((uint32_t)i);
    return !node_proxies.empty();
}

// ---------------------------------------------------------------- UBO (Engine.cpp:642-662, 736-849)
static float HaltonSequence(int index, int base) {
    float f = 1.0f, result = 0.0f;
    while (index > 0) {
        f = f / base;
        result = result + f * (index % base);
        index = index / base;
    }
    return result;
}

GkUniformBufferObject Scene::GetUniformBufferObject(uint32_t width, uint32_t height, uint32_t total_frames, bool progressive,
                                                    float model_rotation_deg) {
    GkUniformBufferObject ubo{};
    mat4 mv = camera.ModelView;
    if (model_rotation_deg != 0.0f) {
        // ModelViewController::ModelView (ModelViewController.cpp:32-42): view * rotateY(modelRotX/300 * 90deg);
        // the Motion benchmark drives modelRotX = t*20 (gkNextMotionBenchmark.cpp:50-51).
        const float cameraRotY = model_rotation_deg / 300.0f;
        mv = mv * rotate(mat4(1.0f), cameraRotY * (90.0f * 0.01745329251994329577f), vec3(0, 1, 0));
    }
    mat4 proj = perspective_rh_zo(camera.FieldOfView * 0.01745329251994329577f, width / static_cast<float>(height), 0.2f, 2000.0f);
    proj[1][1] *= -1;
    mat4 projUnJit = proj;
    if (settings.TAA) {
        int idx = (int)(total_frames % (uint32_t)settings.TemporalFrames);
        float jx = HaltonSequence(idx + 1, 2) - 0.5f, jy = HaltonSequence(idx + 1, 3) - 0.5f;
        proj[2][0] = jx / static_cast<float>(width) * 2.0f;
        proj[2][1] = jy / static_cast<float>(height) * 2.0f;
    }
    mat4 mvi = inverse(mv), pi = inverse(proj), vp = proj * mv, vpu = projUnJit * mv;
    ubo.ModelView = mv.pod();
    ubo.Projection = proj.pod();
    ubo.ModelViewInverse = mvi.pod();
    ubo.ProjectionInverse = pi.pod();
    ubo.ViewProjection = vp.pod();
    ubo.ViewProjectionUnJit = vpu.pod();
    bool prevValid = has_prev_ubo_ && prev_ubo_.TotalFrames != 0;
    ubo.PrevViewProjection = prevValid ? prev_ubo_.ViewProjection : ubo.ViewProjection;
    ubo.PrevViewProjectionUnJit = prevValid ? prev_ubo_.ViewProjectionUnJit : ubo.ViewProjectionUnJit;
    ubo.ViewportRect = {0, 0, (float)width, (float)height};
    ubo.SunViewProjection = mat4(1.0f).pod();  // shadow-map path is not on the PathTracing renderer's path
    ubo.Aperture = camera.Aperture;
    ubo.FocusDistance = camera.FocalDistance;
    ubo.SkyRotation = env.SkyRotation;
    ubo.MaxNumberOfBounces = settings.MaxNumberOfBounces;
    ubo.TotalFrames = total_frames;
    ubo.NumberOfSamples = settings.NumberOfSamples;
    ubo.NumberOfBounces = settings.NumberOfBounces;
    ubo.AdaptiveSample = 0;
    ubo.AdaptiveVariance = 6.0f;
    ubo.AdaptiveSteps = 4;
    ubo.TAA = settings.TAA;
    ubo.RandomSeed = 0;  // rand() in the reference; unused by the hot-path shaders
    vec3 sd = env.SunDirection();
    ubo.SunDirection = {sd.x, sd.y, sd.z, 0.0f};
    ubo.SunColor = {env.SunIntensity, env.SunIntensity, env.SunIntensity, 0};
    ubo.SkyIntensity = env.SkyIntensity;
    ubo.SkyIdx = env.SkyIdx;
    ubo.BackGroundColor = {0.4f * 4.0f * env.SkyIntensity, 0.6f * 4.0f * env.SkyIntensity, 1.0f * 4.0f * env.SkyIntensity, 0};
    ubo.HasSky = env.HasSky;
    ubo.HasSun = env.HasSun && env.SunIntensity > 0;
    ubo.ShowHeatmap = 0;
    ubo.HeatmapScale = 1.0f;
    ubo.UseCheckerBoard = 0;
    ubo.TemporalFrames = progressive ? 256 : settings.TemporalFrames;
    ubo.HDR = settings.HDR;
    ubo.PaperWhiteNit = settings.PaperWhiteNit;
    ubo.LightCount = (uint32_t)lights.size();  // Engine.cpp:830
    ubo.BFSigma = settings.DenoiseSigma;
    ubo.BFSigmaLum = settings.DenoiseSigmaLum;
    ubo.BFSigmaNormal = settings.DenoiseSigmaNormal;
    ubo.BFSize = settings.Denoiser ? settings.DenoiseSize : 0;
    ubo.ShowEdge = 0;
    ubo.ProgressiveRender = progressive;
    ubo.FastGather = settings.FastGather;
    ubo.FastInterpole = 0;
    ubo.DebugDraw_Lighting = settings.DebugDraw_Lighting;
    ubo.DisableSpatialReuse = 0;
    ubo.SuperResolution = settings.SuperResolution;
    prev_ubo_ = ubo;
    has_prev_ubo_ = true;
    return ubo;
}

// ---------------------------------------------------------------- config-4 dynamics
void Scene::JitterInstances(uint32_t step, float fraction) {
    std::mt19937 rng(977u + step);
    uint32_t n = (uint32_t)nodes.size();
    uint32_t moves = (uint32_t)(fraction * n);
    for (uint32_t k = 0; k < moves; ++k) {
        uint32_t i = 1 + rng() % (n - 1);  // never the floor
        Node& nd = nodes[i];
        nd.translation.y += 0.24f * (float)((int)(rng() % 3) - 1);
        if (nd.translation.y < 0) nd.translation.y = 0;
        nd.RecalcTransform();
    }
}

void Scene::PerturbModel(uint32_t model, uint32_t step, std::vector<uint16_t>& out) {
    const Model& m = models[model];
    out.resize(m.vertices.size() * 4);
    float amp = 0.01f * std::sin(0.7f * (float)step);
    for (size_t i = 0; i < m.vertices.size(); ++i) {
        const Vertex& v = m.vertices[i];
        vec3 p = v.Position + v.Normal * amp;
        out[4 * i + 0] = to_float16(p.x); out[4 * i + 1] = to_float16(p.y); out[4 * i + 2] = to_float16(p.z); out[4 * i + 3] = to_float16(p.x);
    }
}

// ---------------------------------------------------------------- procedural sky
GkSphericalHarmonics ProjectHdrToSh(const float* hdrPixels, int width, int height) {
    GkSphericalHarmonics result{};
    const float pi = 3.14159265358979323846f;
    const float shC0 = 0.282095f, shC1 = 0.488603f, shC2 = 1.092548f, shC3 = 0.315392f, shC4 = 0.546274f;
    for (int y = 0; y < height; ++y) {
        const float v = (y + 0.5f) / height;
        const float theta = v * pi;
        const float sinTheta = std::sin(theta), cosTheta = std::cos(theta);
        const float weight = sinTheta * (pi / height) * (2.0f * pi / width);  // solid angle of the texel
        for (int x = 0; x < width; ++x) {
            const float u = (x + 0.5f) / width;
            const float phi = u * 2.0f * pi;
            const float dx = sinTheta * std::cos(phi), dy = cosTheta, dz = sinTheta * std::sin(phi);
            float basis[9];
            basis[0] = shC0;
            basis[1] = -shC1 * dy;
            basis[2] = shC1 * dz;
            basis[3] = -shC1 * dx;
            basis[4] = shC2 * dx * dy;
            basis[5] = -shC2 * dy * dz;
            basis[6] = shC3 * (3.0f * dy * dy - 1.0f);
            basis[7] = -shC2 * dx * dz;
            basis[8] = shC4 * (dx * dx - dz * dz);
            const float* px = hdrPixels + ((size_t)y * width + x) * 4;
            for (int i = 0; i < 9; ++i) {
                result.coefficients[0][i] += px[0] * basis[i] * weight;
                result.coefficients[1][i] += px[1] * basis[i] * weight;
                result.coefficients[2][i] += px[2] * basis[i] * weight;
            }
        }
    }
    return result;
}

void MakeProceduralSky(uint32_t width, uint32_t height, std::vector<float>& rgba) {
    rgba.resize((size_t)width * height * 4);
    const float pi = 3.14159265358979323846f;
    const vec3 sun = normalize(vec3(0.4f, 0.55f, -0.73f));
    for (uint32_t y = 0; y < height; ++y) {
        float theta = (y + 0.5f) / height * pi;  // acos(d.y)
        for (uint32_t x = 0; x < width; ++x) {
            float phi = (x + 0.5f) / width * 2 * pi;  // atan2(d.x, d.z)
            vec3 d(std::sin(theta) * std::sin(phi), std::cos(theta), std::sin(theta) * std::cos(phi));
            float up = d.y;
            vec3 c;
            if (up >= 0) {
                float t = std::pow(1.0f - up, 3.0f);
                c = vec3(0.25f, 0.45f, 0.9f) * (1 - t) + vec3(0.9f, 0.85f, 0.8f) * t;
            } else {
                float t = std::min(1.0f, -up * 4.0f);
                c = vec3(0.9f, 0.85f, 0.8f) * (1 - t) + vec3(0.18f, 0.16f, 0.14f) * t;
            }
            float cs = std::max(0.0f, dot(d, sun));
            c = c + vec3(1.0f, 0.9f, 0.7f) * (8.0f * std::pow(cs, 400.0f) + 0.4f * std::pow(cs, 8.0f));
            size_t o = ((size_t)y * width + x) * 4;
            rgba[o] = c.x; rgba[o + 1] = c.y; rgba[o + 2] = c.z; rgba[o + 3] = 1.0f;
        }
    }
}

// The procedural generators by number (0 Cornell box: args = scale; 1 box: p0, p1; 2 sphere: centre, radius), for tools and tests.
Model CreateProcModel(int kind, const float* a, std::vector<GkMaterial>& materials, std::vector<GkLightObject>& lights) {
    if (kind == 0) return CreateCornellBox(a[0], materials, lights);
    if (kind == 1) return CreateBox(vec3(a[0], a[1], a[2]), vec3(a[3], a[4], a[5]));
    return CreateSphere(vec3(a[0], a[1], a[2]), a[3]);
}
GkMaterial MakeMaterialByKind(int kind, vec3 d, float a, float b) {  // Material.hpp:10-38 in declaration order
    switch (kind) {
        case 0: return MaterialLambertian(d);
        case 1: return MaterialMetallic(d, a);
        case 2: return MaterialMixture(d, a);
        case 3: return MaterialDielectric(a, b);
        default: return MaterialDiffuseLight(d);
    }
}
}  // namespace gk
