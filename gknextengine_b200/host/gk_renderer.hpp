// C++ host-side mirror of the reference's renderer plug-in interface for the path-tracing hot path, sitting directly above the
// C ABI (include/gkcu.h). Same names, argument meaning and error behaviour as the reference so that an engine-side adaptor is
// a thin forwarder:
//
//   reference                                                        here
//   ---------------------------------------------------------------  ------------------------------------------------
//   enum ERendererType                 VulkanBaseRenderer.hpp:45-51   gk::ERendererType (+ ERT_CudaPathTracing)
//   class LogicRendererBase            VulkanBaseRenderer.hpp:213-243 gk::LogicRendererBase
//   RayTracing::PathTracingRenderer    PathTracingRenderer.hpp:30-71  gk::CudaPathTracingRenderer
//   RegisterLogicRenderer / Switch...  VulkanBaseRenderer.cpp:892-917 gk::HeadlessBaseRenderer
//   RayTraceBaseRenderer AS hooks      RayTraceBaseRenderer.cpp:165-218 OnPreLoadScene / OnPostLoadScene / AfterUpdateScene / PreRender
//   RayTraceBaseRenderer::PostRender   RayTraceBaseRenderer.cpp:225-274 CudaPathTracingRenderer::PostRender (probe light bake)
//   Throw(std::runtime_error)          src/Utilities/Exception.hpp:9-15  gk::Check() throws std::runtime_error with gkcu_last_error()
//
// The Vulkan-typed parameters of the reference (VkExtent2D, VkCommandBuffer, imageIndex) become plain values: the CUDA
// renderer records nothing into a Vulkan command buffer; it enqueues on its own stream.
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/gkcu.h"
#include "gk_scene.hpp"

namespace gk {

enum ERendererType {
    ERT_PathTracing,
    ERT_ModernDeferred,
    ERT_LegacyDeferred,
    ERT_VoxelTracing,
    ERT_CudaPathTracing,  // new: the B200 renderer behind the same switch
};

struct Extent2D { uint32_t width, height; };

inline void Check(int rc, const char* operation) {  // src/Vulkan/Vulkan.cpp:7-13 analogue
    if (rc != 0) throw std::runtime_error(std::string("failed to ") + operation + " (" + gkcu_last_error() + ")");
}

class HeadlessBaseRenderer;

class LogicRendererBase {
public:
    explicit LogicRendererBase(HeadlessBaseRenderer& baseRender) : baseRender_(baseRender) {}
    virtual ~LogicRendererBase() {}
    virtual void OnDeviceSet() {}
    virtual void CreateSwapChain(const Extent2D& extent) {}
    virtual void DeleteSwapChain() {}
    virtual void Render(const GkUniformBufferObject& ubo, uint32_t imageIndex) {}
    virtual void BeforeNextFrame() {}
    HeadlessBaseRenderer& baseRender_;
};

// The part of VulkanBaseRenderer / RayTraceBaseRenderer the path needs, without a window or a swap chain.
class HeadlessBaseRenderer {
public:
    explicit HeadlessBaseRenderer(int device = 0) : device_(device) {}
    virtual ~HeadlessBaseRenderer() {}

    void RegisterLogicRenderer(ERendererType type);
    void SwitchLogicRenderer(ERendererType type) { currentLogicRenderer_ = type; }
    ERendererType CurrentLogicRendererType() const { return currentLogicRenderer_; }
    LogicRendererBase* CurrentLogicRenderer() { auto it = logicRenderers_.find(currentLogicRenderer_); return it == logicRenderers_.end() ? nullptr : it->second.get(); }

    void SetScene(Scene* scene) { scene_ = scene; }
    Scene& GetScene() { return *scene_; }
    int Device() const { return device_; }
    int FrameCount() const { return frameCount_; }

    void CreateSwapChain(const Extent2D& extent);
    void DeleteSwapChain();
    void OnPreLoadScene();    // frees the acceleration structures (RayTraceBaseRenderer.cpp:193-197)
    void OnPostLoadScene();   // uploads the mesh buffers and builds BLAS + TLAS (:199-203, :106-124)
    void AfterUpdateScene();  // re-uploads NodeProxy[] / instance records (:165-191)
    void DrawFrame();         // UpdateNodes -> AfterUpdateScene -> PreRender -> Render of the current logic renderer

    bool progressiveRendering_ = false;
    bool bakeProbes_ = false;   // run the probe voxelisation at scene load and the light bake every frame
    int bakeSpeedLevel_ = 1;    // UserSettings.BakeSpeedLevel: 0 -> 30 frames per sweep, 1 -> 120, 2 -> off
    std::vector<float> sky_;  // RGBA32F equirect for sky id 0
    uint32_t skyW_ = 0, skyH_ = 0;

private:
    friend class CudaPathTracingRenderer;
    int device_;
    Scene* scene_ = nullptr;
    Extent2D extent_{0, 0};
    uint32_t frameCount_ = 0;
    std::map<ERendererType, std::unique_ptr<LogicRendererBase>> logicRenderers_;
    ERendererType currentLogicRenderer_ = ERT_CudaPathTracing;
};

class CudaPathTracingRenderer final : public LogicRendererBase {
public:
    explicit CudaPathTracingRenderer(HeadlessBaseRenderer& baseRender);
    ~CudaPathTracingRenderer() override;
    void OnDeviceSet() override;
    void CreateSwapChain(const Extent2D& extent) override;
    void DeleteSwapChain() override;
    void Render(const GkUniformBufferObject& ubo, uint32_t imageIndex) override;
    void BeforeNextFrame() override;

    // acceleration-structure hooks forwarded from the base renderer
    void OnPreLoadScene();
    void OnPostLoadScene(Scene& scene, const float* sky, uint32_t skyW, uint32_t skyH);
    void AfterUpdateScene(Scene& scene);
    void PreRender();
    // RayTraceBaseRenderer::PostRender (RayTraceBaseRenderer.cpp:225-274): the "hw-lightbake" dispatch over 1/temporalFrames of the grid
    void PostRender(const GkUniformBufferObject& ubo, uint32_t frameCount);
    // FCPUAccelerationStructure::AsyncProcessFull(incremental = false) + Tick upload, done in one call on the GPU
    void VoxelizeProbes();

    void Readback(uint32_t rtIndex, void* dst, size_t bytes);  // RT_DENOISED is what the base renderer blits to the swap chain
    GkcuTimings Timings();
    GkcuCounters Counters();
    GkcuContext* Handle() { return ctx_; }

private:
    GkcuContext* ctx_ = nullptr;
    bool sceneLoaded_ = false;
};

}  // namespace gk
