"""Bench workload with the probe grid active (voxelised + one full bake sweep): python scripts/profile_probes.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gknextengine_b200 import CudaPathTracingRenderer, GkcuFrameParams, host
W, H = 1920, 1080
scene = host.Scene("RTIO.proc", samples=16, bounces=8, taa=True, temporal_frames=16, denoiser=True)
scene.update_nodes()
sky = host.make_sky(1024, 512)
r = CudaPathTracingRenderer(0)
r.set_kernel_timing(True)  # per-launch event brackets: kernel_stats() needs them
r.CreateSwapChain((W, H))
r.OnPostLoadScene(scene, sky)
r.set_lights(scene); r.set_sky_sh(0, host.project_sh(sky))
p = GkcuFrameParams(temporal_frames=16, spp_world=1)
modes = (True,) if "--probes-only" in sys.argv else (False, True)
frames = 2 if "--probes-only" in sys.argv else 5
for probes in modes:
    if probes:
        r.probes_voxelize()
        ubo = scene.ubo(W, H, 1)
        for f in range(120): r.PostRender(ubo, f)
    for f in range(1, frames):
        r.PreRender(); r.Render(scene.ubo(W, H, f), p); r.BeforeNextFrame()
    t, k = r.timings(), r.kernel_stats()
    print("probes" if probes else "zero  ", {x: round(t[x], 2) for x in ("rt_pass_ms",)}, {x: round(k[x], 2) for x in ("trace_ms", "shade_ms", "tail_ms", "primary_ms")}, r.counters()["closest_hit_rays"])
r.close()
