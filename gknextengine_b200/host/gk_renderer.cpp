// See gk_renderer.hpp. Thin forwarding onto the C ABI; errors surface as std::runtime_error like the reference's Throw().
#include "gk_renderer.hpp"

namespace gk {

void HeadlessBaseRenderer::RegisterLogicRenderer(ERendererType type) {  // VulkanBaseRenderer.cpp:892-912
    switch (type) {
        case ERT_CudaPathTracing:
            logicRenderers_[type] = std::make_unique<CudaPathTracingRenderer>(*this);
            break;
        default:
            throw std::runtime_error("only ERT_CudaPathTracing is part of this build; the other logic renderers are Vulkan renderers of the reference");
    }
    currentLogicRenderer_ = type;
    logicRenderers_[type]->OnDeviceSet();
}

void HeadlessBaseRenderer::CreateSwapChain(const Extent2D& extent) {  // VulkanBaseRenderer.cpp:465-524
    extent_ = extent;
    for (auto& lr : logicRenderers_) lr.second->CreateSwapChain(extent);
}

void HeadlessBaseRenderer::DeleteSwapChain() {
    for (auto& lr : logicRenderers_) lr.second->DeleteSwapChain();
}

void HeadlessBaseRenderer::OnPreLoadScene() {
    for (auto& lr : logicRenderers_)
        if (auto* c = dynamic_cast<CudaPathTracingRenderer*>(lr.second.get())) c->OnPreLoadScene();
}

void HeadlessBaseRenderer::OnPostLoadScene() {
    frameCount_ = 0;
    for (auto& lr : logicRenderers_)
        if (auto* c = dynamic_cast<CudaPathTracingRenderer*>(lr.second.get())) c->OnPostLoadScene(*scene_, sky_.empty() ? nullptr : sky_.data(), skyW_, skyH_);
}

void HeadlessBaseRenderer::AfterUpdateScene() {
    for (auto& lr : logicRenderers_)
        if (auto* c = dynamic_cast<CudaPathTracingRenderer*>(lr.second.get())) c->AfterUpdateScene(*scene_);
}

void HeadlessBaseRenderer::DrawFrame() {  // VulkanBaseRenderer.cpp:727-869, without acquire/present
    auto it = logicRenderers_.find(currentLogicRenderer_);
    if (it == logicRenderers_.end()) throw std::runtime_error("no logic renderer registered");
    it->second->BeforeNextFrame();
    if (scene_->UpdateNodes()) AfterUpdateScene();  // :806-809
    if (auto* c = dynamic_cast<CudaPathTracingRenderer*>(it->second.get())) c->PreRender();
    GkUniformBufferObject ubo = scene_->GetUniformBufferObject(extent_.width, extent_.height, frameCount_, progressiveRendering_);
    it->second->Render(ubo, 0);
    if (bakeProbes_)
        if (auto* c = dynamic_cast<CudaPathTracingRenderer*>(it->second.get())) c->PostRender(ubo, frameCount_);
    ++frameCount_;
}

CudaPathTracingRenderer::CudaPathTracingRenderer(HeadlessBaseRenderer& baseRender) : LogicRendererBase(baseRender) {}

CudaPathTracingRenderer::~CudaPathTracingRenderer() {
    if (ctx_) gkcu_destroy(ctx_);
}

void CudaPathTracingRenderer::OnDeviceSet() {
    if (!ctx_) Check(gkcu_create(baseRender_.Device(), &ctx_), "create the CUDA path tracing context");
}

void CudaPathTracingRenderer::CreateSwapChain(const Extent2D& extent) {  // PathTracingRenderer.cpp:66-90
    Check(gkcu_resize(ctx_, extent.width, extent.height), "create the storage images");
}

void CudaPathTracingRenderer::DeleteSwapChain() {}

void CudaPathTracingRenderer::OnPreLoadScene() { sceneLoaded_ = false; }

void CudaPathTracingRenderer::OnPostLoadScene(Scene& scene, const float* sky, uint32_t skyW, uint32_t skyH) {
    GkcuSceneDesc d{};
    d.vertices = scene.gpu_vertices.data();
    d.vertex_count = (uint32_t)scene.gpu_vertices.size();
    d.simple_vertices = scene.simple_vertices.data();
    d.indices = scene.prim_indices.data();
    d.index_count = (uint32_t)scene.prim_indices.size();
    d.offsets = scene.offsets.data();
    d.offset_count = (uint32_t)scene.offsets.size();
    d.model_count = (uint32_t)scene.models.size();
    Check(gkcu_scene_upload(ctx_, &d), "upload the mesh buffers");
    Check(gkcu_update_materials(ctx_, scene.materials.data(), (uint32_t)scene.materials.size()), "upload the materials");
    Check(gkcu_blas_build(ctx_), "build the bottom level acceleration structures");
    if (sky) {
        Check(gkcu_set_sky(ctx_, 0, sky, skyW, skyH), "upload the sky");
        const GkSphericalHarmonics sh = ProjectHdrToSh(sky, (int)skyW, (int)skyH);  // Texture.cpp:448-449
        Check(gkcu_set_sky_sh(ctx_, 0, &sh), "upload the sky SH");
    }
    Check(gkcu_set_lights(ctx_, scene.lights.data(), (uint32_t)scene.lights.size()), "upload the lights");
    AfterUpdateScene(scene);
    sceneLoaded_ = true;
    if (baseRender_.bakeProbes_) VoxelizeProbes();
}

void CudaPathTracingRenderer::AfterUpdateScene(Scene& scene) {
    Check(gkcu_update_nodes(ctx_, scene.node_proxies.data(), (uint32_t)scene.node_proxies.size()), "upload the node proxies");
}

void CudaPathTracingRenderer::PreRender() { Check(gkcu_tlas_build(ctx_), "rebuild the top level acceleration structure"); }

void CudaPathTracingRenderer::Render(const GkUniformBufferObject& ubo, uint32_t) {  // PathTracingRenderer.cpp:104-218
    GkcuFrameParams p{};
    p.progressive = baseRender_.progressiveRendering_ ? 1 : 0;
    p.temporal_frames = baseRender_.GetScene().settings.TemporalFrames;
    p.spp_world = 1;
    Check(gkcu_render_frame(ctx_, &ubo, &p), "render the frame");
}

void CudaPathTracingRenderer::PostRender(const GkUniformBufferObject& ubo, uint32_t frameCount) {
    int temporalFrames;
    switch (baseRender_.bakeSpeedLevel_) {
        case 0: temporalFrames = 30; break;
        case 2: return;  // BakeSpeedLevel 2 records no dispatch
        default: temporalFrames = 120; break;
    }
    const int cubesPerGroup = 64;
    const int group = GK_CUBE_SIZE_XY * GK_CUBE_SIZE_XY * GK_CUBE_SIZE_Z / cubesPerGroup;
    const int groupPerFrame = group / temporalFrames;
    const int offsetInCubes = (int)(frameCount % temporalFrames) * groupPerFrame * cubesPerGroup;
    Check(gkcu_probes_bake(ctx_, &ubo, (uint32_t)offsetInCubes, (uint32_t)(groupPerFrame * cubesPerGroup)), "bake the ambient cubes");
}

void CudaPathTracingRenderer::VoxelizeProbes() { Check(gkcu_probes_voxelize(ctx_), "voxelize the probe grid"); }

void CudaPathTracingRenderer::BeforeNextFrame() { Check(gkcu_synchronize(ctx_), "wait for the previous frame"); }

void CudaPathTracingRenderer::Readback(uint32_t rtIndex, void* dst, size_t bytes) { Check(gkcu_readback(ctx_, rtIndex, dst, bytes), "read back a storage image"); }

GkcuTimings CudaPathTracingRenderer::Timings() {
    GkcuTimings t{};
    Check(gkcu_get_timings(ctx_, &t), "query timings");
    return t;
}

GkcuCounters CudaPathTracingRenderer::Counters() {
    GkcuCounters c{};
    Check(gkcu_get_counters(ctx_, &c), "query counters");
    return c;
}

}  // namespace gk
