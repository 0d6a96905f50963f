// Headless C++ harness: plays the role of the engine (NextEngine::Tick -> renderer->DrawFrame) for the CUDA renderer.
//   gk_headless <scene> <width> <height> <frames> [spp] [bounces] [bake]     bake = 1: voxelise the probe grid at load, bake 1/120 per frame
// Prints per-pass GPU times, exact ray counts and writes RT_DENOISED of the last frame as a PPM.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

#include "gk_renderer.hpp"

static float half_to_float(uint16_t h) {
    uint32_t s = (h >> 15) & 1, e = (h >> 10) & 31, m = h & 1023;
    if (e == 0) return (s ? -1.f : 1.f) * std::ldexp((float)m, -24);
    if (e == 31) return m ? NAN : (s ? -INFINITY : INFINITY);
    return (s ? -1.f : 1.f) * std::ldexp((float)(m | 1024), (int)e - 25);
}

int main(int argc, char** argv) {
    const char* name = argc > 1 ? argv[1] : "CornellBox.proc";
    uint32_t w = argc > 2 ? atoi(argv[2]) : 640, h = argc > 3 ? atoi(argv[3]) : 360;
    int frames = argc > 4 ? atoi(argv[4]) : 8;
    try {
        gk::Scene scene;
        if (!scene.Load(name)) { std::fprintf(stderr, "unknown scene %s\n", name); return 2; }
        scene.settings.NumberOfSamples = argc > 5 ? atoi(argv[5]) : 1;
        scene.settings.NumberOfBounces = argc > 6 ? atoi(argv[6]) : 4;
        scene.settings.Denoiser = true;
        gk::HeadlessBaseRenderer base(0);
        base.SetScene(&scene);
        gk::MakeProceduralSky(1024, 512, base.sky_);
        base.skyW_ = 1024;
        base.skyH_ = 512;
        base.bakeProbes_ = argc > 7 && atoi(argv[7]) != 0;
        base.RegisterLogicRenderer(gk::ERT_CudaPathTracing);  // the renderer switch: Engine.cpp:57-90 / gkNextRenderer.cpp:316-321
        base.SwitchLogicRenderer(gk::ERT_CudaPathTracing);
        base.CreateSwapChain({w, h});
        base.OnPreLoadScene();
        base.OnPostLoadScene();
        auto* r = static_cast<gk::CudaPathTracingRenderer*>(base.CurrentLogicRenderer());
        for (int f = 0; f < frames; ++f) {
            base.DrawFrame();
            r->BeforeNextFrame();
            GkcuTimings t = r->Timings();
            GkcuCounters c = r->Counters();
            double rays = (double)(c.primary_rays + c.closest_hit_rays + c.any_hit_rays);
            std::printf("frame %d: TLAS Update %.3f ms | rt pass %.3f ms | reproject pass %.3f ms | compose pass %.3f ms | frame %.3f ms | %.0f rays, %.1f Mrays/s\n",
                        f, t.tlas_update_ms, t.rt_pass_ms, t.reproject_pass_ms, t.compose_pass_ms, t.frame_ms, rays, rays / (t.rt_pass_ms * 1e3));
        }
        std::vector<uint16_t> img((size_t)w * h * 4);
        r->Readback(GK_RT_DENOISED, img.data(), img.size() * 2);
        if (FILE* fp = std::fopen("gk_headless.ppm", "wb")) {
            std::fprintf(fp, "P6\n%u %u\n255\n", w, h);
            for (size_t i = 0; i < (size_t)w * h; ++i)
                for (int k = 0; k < 3; ++k) {
                    float v = half_to_float(img[i * 4 + k]);
                    v = std::pow(std::fmin(std::fmax(v, 0.f), 1.f), 1.f / 2.2f);
                    std::fputc((int)(v * 255.f + 0.5f), fp);
                }
            std::fclose(fp);
        }
    } catch (const std::exception& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
