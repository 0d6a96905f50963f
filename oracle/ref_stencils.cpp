// Compiles the reference's two stencil compute shaders as C++ and runs them over caller-provided storage images, so that the oracle's
// restatements of SURVEY 8a rows a23 / a24 can be pinned against the reference's own text:
//   assets/shaders/Process.ReProject.comp.slang   (whole file: calculateWeight, FilterHistoryColor, main)
//   assets/shaders/Process.DenoiseJBF.comp.slang  (whole file: the groupshared tile, JBF, main)
// oracle/Makefile copies each file through sed into oracle/_ref/*_gen.inc where it only rewrites what C++ cannot parse: the
// [shader]/[numthreads] attributes, `import`, the `: SV_*` semantics, `let` -> `auto`, `groupshared` -> `static`,
// `inout T x` -> `T& x`, the push-constant declaration, `main` -> a unique name, and the one reversed swizzle `.yx`. The bodies are the
// reference's. This file supplies the vocabulary: HLSL vectors (ref_hlsl_shim.h), storage images in the reference's formats
// (VulkanBaseRenderer.cpp:439-457: fp16x4 with round-to-nearest-even stores, R32_UINT, RG32F, R16_UINT; out-of-range loads return 0
// and out-of-range stores are dropped, i.e. robust image access), the camera block (the reference's own BasicTypes.slang compiled as
// C++), and a thread group: DenoiseJBF's 8 x 8 group runs on 64 threads that meet at GroupMemoryBarrierWithGroupSync().
// The tonemappers come from the reference's Const_Func.slang extract (const_func_gen.inc). Built only when /root/reference is mounted.
// TEST INFRASTRUCTURE ONLY.
#include <barrier>
#include <thread>
#include <vector>

#define REF_DET_MATH
#include "ref_image_shim.h"

namespace ref {
// ---- a thread group for GroupMemoryBarrierWithGroupSync()
static std::barrier<>* g_group_barrier = nullptr;
inline void GroupMemoryBarrierWithGroupSync() { g_group_barrier->arrive_and_wait(); }

#include "PreProcessor.slang"
#include "const_func_gen.inc"  // GT_Tonemapping, LinearToST2084UE (and the rest of the Const_Func.slang extract)
namespace reproject {
#include "reproject_gen.inc"
}
namespace denoise {
#include "denoise_gen.inc"
}
}  // namespace ref

using namespace ref;

// images[rt] = pointer to the storage image in its reference format (or NULL), indexed by the reference's RT_* numbers
static void bind(void* const* images, int w, int h) {
    const int f16[] = {0, 1, 6, 7, 9, 11, 12, 13, 20, 21, 22};
    for (int i : f16) Bindless.f4[i] = ImgF4{(uint8_t*)images[i], w, h, FMT_F16X4};
    Bindless.f4[5] = ImgF4{(uint8_t*)images[5], w, h, FMT_RG32F};
    Bindless.u[3] = ImgU{(uint8_t*)images[3], w, h, 4};
    Bindless.u[4] = ImgU{(uint8_t*)images[4], w, h, 4};
    Bindless.u[14] = ImgU{(uint8_t*)images[14], w, h, 2};
}

extern "C" {
// PathTracingRenderer::Render's reproject dispatch (PathTracingRenderer.cpp:126-136): ceil(w/8) x ceil(h/8) groups of 8 x 8
void ref_reproject(void* const* images, int w, int h, int progressive, int temporal_frames, int fast_reproject) {
    bind(images, w, h);
    reproject::pushConsts.ProgressiveRender = progressive;
    reproject::pushConsts.TemporalFrames = temporal_frames;
    reproject::pushConsts.PrevDiffuseBindlessIdx = BindlessShim::RT_SINGLE_PREV_DIFFUSE;
    reproject::pushConsts.PrevSpecularBindlessIdx = BindlessShim::RT_SINGLE_PREV_SPECULAR;
    reproject::pushConsts.PrevAlbedoBindlessIdx = BindlessShim::RT_SINGLE_PREV_ALBEDO;
    reproject::pushConsts.FastReproject = fast_reproject;
    for (int y = 0; y < (h + 7) / 8 * 8; ++y)
        for (int x = 0; x < (w + 7) / 8 * 8; ++x) reproject::reproject_main(uint3((uint)x, (uint)y, 0u));
}

// the compose dispatch (PathTracingRenderer.cpp:138-157); ubo = the 784-byte UniformBufferObject
void ref_denoise(void* const* images, int w, int h, const void* ubo) {
    bind(images, w, h);
    UniformBufferObject cam;
    static_assert(sizeof(UniformBufferObject) == 784, "UniformBufferObject");
    std::memcpy(&cam, ubo, sizeof(cam));
    Bindless.scene.Camera = &cam;
    for (int gy = 0; gy < (h + 7) / 8; ++gy)
        for (int gx = 0; gx < (w + 7) / 8; ++gx) {
            std::barrier<> bar(64);
            g_group_barrier = &bar;
            std::vector<std::thread> pool;
            for (int t = 0; t < 64; ++t)
                pool.emplace_back([=]() {
                    const uint tx = t & 7, ty = t >> 3;
                    denoise::denoise_main(uint3(gx * 8 + tx, gy * 8 + ty, 0u), uint3(tx, ty, 0u), uint3((uint)gx, (uint)gy, 0u), (uint)t);
                });
            for (auto& th : pool) th.join();
        }
    Bindless.scene.Camera = nullptr;
}
}
