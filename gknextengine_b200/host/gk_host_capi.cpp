// C ABI over the C++ host-side scene code (gk_scene.hpp) so the Python tests / bench can obtain exactly the arrays the
// reference's Scene hands to its renderer: libgknext_host.so. Pure CPU; no CUDA here.
#include <cstring>
#include <string>

#include "gk_scene.hpp"

static thread_local std::string g_host_error;

extern "C" {

const char* gkhost_last_error() { return g_host_error.c_str(); }

void* gkhost_scene_create(const char* name) {
    auto* s = new gk::Scene();
    if (!s->Load(name)) {
        g_host_error = s->last_error.empty() ? std::string("unknown scene ") + name : s->last_error;
        delete s;
        return nullptr;
    }
    return s;
}
void gkhost_scene_destroy(void* h) { delete static_cast<gk::Scene*>(h); }

void gkhost_scene_desc(void* h, GkcuSceneDesc* out) {
    auto* s = static_cast<gk::Scene*>(h);
    out->vertices = s->gpu_vertices.data();
    out->vertex_count = (uint32_t)s->gpu_vertices.size();
    out->simple_vertices = s->simple_vertices.data();
    out->indices = s->prim_indices.data();
    out->index_count = (uint32_t)s->prim_indices.size();
    out->offsets = s->offsets.data();
    out->offset_count = (uint32_t)s->offsets.size();
    out->model_count = (uint32_t)s->models.size();
}
// The raster-side buffers of Scene::RebuildMeshBuffer (Scene.cpp:224-250): the provoking index buffer and GPUScene.Reorders, plus the
// authored (pre-rotation) model-local indices of model `model` for checks.
const uint32_t* gkhost_scene_provoke_indices(void* h, uint32_t* count) {
    auto* s = static_cast<gk::Scene*>(h);
    *count = (uint32_t)s->provoke_indices.size();
    return s->provoke_indices.data();
}
const uint32_t* gkhost_scene_reorders(void* h, uint32_t* count) {
    auto* s = static_cast<gk::Scene*>(h);
    *count = (uint32_t)s->reorders.size();
    return s->reorders.data();
}
const uint32_t* gkhost_scene_model_indices(void* h, uint32_t model, uint32_t* count) {
    auto* s = static_cast<gk::Scene*>(h);
    if (model >= s->models.size()) { *count = 0; return nullptr; }
    *count = (uint32_t)s->models[model].indices.size();
    return s->models[model].indices.data();
}

// Radiance .hdr -> RGBA32F. First call with rgba = NULL for the size; returns 0 on success, the message through gkhost_last_error().
int gkhost_load_hdr(const char* path, float* rgba, uint32_t* width, uint32_t* height) {
    std::vector<float> px;
    uint32_t w = 0, h = 0;
    if (!gk::LoadHdrImage(path, px, w, h, g_host_error)) return 1;
    *width = w;
    *height = h;
    if (rgba) std::memcpy(rgba, px.data(), px.size() * sizeof(float));
    return 0;
}

uint32_t gkhost_provoking_index_buffer(uint32_t* destination, uint32_t* reorder, const uint32_t* indices, uint32_t index_count, uint32_t vertex_count) {
    return (uint32_t)gk::GenerateProvokingIndexBuffer(destination, reorder, indices, index_count, vertex_count);
}

const GkNodeProxy* gkhost_scene_nodes(void* h, uint32_t* count) {
    auto* s = static_cast<gk::Scene*>(h);
    *count = (uint32_t)s->node_proxies.size();
    return s->node_proxies.data();
}
const GkMaterial* gkhost_scene_materials(void* h, uint32_t* count) {
    auto* s = static_cast<gk::Scene*>(h);
    *count = (uint32_t)s->materials.size();
    return s->materials.data();
}
const GkLightObject* gkhost_scene_lights(void* h, uint32_t* count) {
    auto* s = static_cast<gk::Scene*>(h);
    *count = (uint32_t)s->lights.size();
    return s->lights.data();
}
void gkhost_project_sh(const float* rgba, uint32_t width, uint32_t height, GkSphericalHarmonics* out) {
    *out = gk::ProjectHdrToSh(rgba, (int)width, (int)height);
}
// proxies changed by the last gkhost_scene_update_nodes; *all != 0: upload everything
const uint32_t* gkhost_scene_changed(void* h, uint32_t* count, int* all) {
    auto* s = static_cast<gk::Scene*>(h);
    *count = (uint32_t)s->changed_proxies.size();
    *all = s->all_proxies_changed ? 1 : 0;
    return s->changed_proxies.data();
}
// Scene::UpdateNodes: ticks node velocities (first call after load reports every node as moved, Node.cpp:154-171)
int gkhost_scene_update_nodes(void* h) { return static_cast<gk::Scene*>(h)->UpdateNodes() ? 1 : 0; }

void gkhost_scene_settings(void* h, int samples, int bounces, int max_bounces, int taa, int temporal_frames, int denoiser, int hdr,
                           int fast_gather) {
    auto* s = static_cast<gk::Scene*>(h);
    s->settings.NumberOfSamples = samples;
    s->settings.NumberOfBounces = bounces;
    s->settings.MaxNumberOfBounces = max_bounces;
    s->settings.TAA = taa != 0;
    s->settings.TemporalFrames = temporal_frames;
    s->settings.Denoiser = denoiser != 0;
    s->settings.HDR = hdr != 0;
    s->settings.FastGather = fast_gather != 0;
}
void gkhost_scene_env(void* h, int has_sky, int has_sun, float sky_intensity, float sun_intensity, float sun_rotation, float sky_rotation) {
    auto* s = static_cast<gk::Scene*>(h);
    s->env.HasSky = has_sky != 0;
    s->env.HasSun = has_sun != 0;
    s->env.SkyIntensity = sky_intensity;
    s->env.SunIntensity = sun_intensity;
    s->env.SunRotation = sun_rotation;
    s->env.SkyRotation = sky_rotation;
}
void gkhost_scene_camera(void* h, float aperture, float focal_distance, float fov) {
    auto* s = static_cast<gk::Scene*>(h);
    s->camera.Aperture = aperture;
    s->camera.FocalDistance = focal_distance;
    if (fov > 0) s->camera.FieldOfView = fov;
}
void gkhost_scene_ubo(void* h, uint32_t width, uint32_t height, uint32_t total_frames, int progressive, float model_rotation_deg,
                      GkUniformBufferObject* out) {
    *out = static_cast<gk::Scene*>(h)->GetUniformBufferObject(width, height, total_frames, progressive != 0, model_rotation_deg);
}
void gkhost_scene_reset_history(void* h) { static_cast<gk::Scene*>(h)->ResetHistory(); }

void gkhost_scene_counts(void* h, uint32_t* nodes, uint32_t* models, uint32_t* materials, uint64_t* instanced_triangles) {
    auto* s = static_cast<gk::Scene*>(h);
    *nodes = (uint32_t)s->nodes.size();
    *models = (uint32_t)s->models.size();
    *materials = (uint32_t)s->materials.size();
    uint64_t t = 0;
    for (auto& n : s->nodes)
        if (n.modelId < s->models.size()) t += s->models[n.modelId].indices.size() / 3;  // camera / group nodes carry no model
    *instanced_triangles = t;
}
// Node::SetVisible (Node.h:41). The next UpdateNodes keeps the node's proxies and clears their `visible` word (Scene.cpp:456-476).
void gkhost_scene_set_node_visible(void* h, uint32_t node, int visible) {
    auto* s = static_cast<gk::Scene*>(h);
    if (node < s->nodes.size()) s->nodes[node].visible = visible != 0;
}
void gkhost_scene_jitter_instances(void* h, uint32_t step, float fraction) { static_cast<gk::Scene*>(h)->JitterInstances(step, fraction); }
// returns the vertex count; out must hold 4 * vertex_count halfs (call with out = NULL to query)
uint32_t gkhost_scene_perturb_model(void* h, uint32_t model, uint32_t step, uint16_t* out) {
    auto* s = static_cast<gk::Scene*>(h);
    uint32_t n = (uint32_t)s->models[model].vertices.size();
    if (!out) return n;
    std::vector<uint16_t> v;
    s->PerturbModel(model, step, v);
    std::memcpy(out, v.data(), v.size() * 2);
    return n;
}
// FProcModel::CreateCornellBox / CreateBox / CreateSphere by number (gk::CreateProcModel). vertices: 13 words each (position 3, normal 3,
// tangent 4, uv 2 as floats, material slot as uint32); pass NULL arrays to query the counts. Returns the number of materials pushed.
uint32_t gkhost_proc_model(int kind, const float* args, void* vertices, uint32_t* vertex_count, uint32_t* indices, uint32_t* index_count,
                           GkMaterial* materials, GkLightObject* light) {
    std::vector<GkMaterial> mats;
    std::vector<GkLightObject> lights;
    gk::Model m = gk::CreateProcModel(kind, args, mats, lights);
    *vertex_count = (uint32_t)m.vertices.size();
    *index_count = (uint32_t)m.indices.size();
    if (vertices) {
        uint32_t* w = static_cast<uint32_t*>(vertices);
        for (const auto& v : m.vertices) {
            const float f[12] = {v.Position.x, v.Position.y, v.Position.z, v.Normal.x, v.Normal.y, v.Normal.z, v.Tangent.x, v.Tangent.y, v.Tangent.z, v.Tangent.w,
                                 v.TexCoord.x, v.TexCoord.y};
            std::memcpy(w, f, 48);
            w[12] = v.MaterialIndex;
            w += 13;
        }
    }
    if (indices) std::memcpy(indices, m.indices.data(), m.indices.size() * 4);
    if (materials) std::memcpy(materials, mats.data(), mats.size() * sizeof(GkMaterial));
    if (light && !lights.empty()) *light = lights[0];
    return (uint32_t)mats.size();
}
void gkhost_material(int kind, const float* diffuse, float a, float b, GkMaterial* out) {
    *out = gk::MakeMaterialByKind(kind, gk::vec3(diffuse[0], diffuse[1], diffuse[2]), a, b);
}
void gkhost_make_sky(uint32_t width, uint32_t height, float* rgba) {
    std::vector<float> v;
    gk::MakeProceduralSky(width, height, v);
    std::memcpy(rgba, v.data(), v.size() * 4);
}
uint16_t gkhost_to_float16(float f) { return gk::to_float16(f); }
void gkhost_mat4_inverse(const float* in, float* out) {
    gk::mat4 m;
    std::memcpy(m.m, in, 64);
    gk::mat4 r = gk::inverse(m);
    std::memcpy(out, r.m, 64);
}

}  // extern "C"
