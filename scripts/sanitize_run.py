"""Small end-to-end exercise of every kernel family for compute-sanitizer:
   compute-sanitizer --tool memcheck|racecheck|initcheck|synccheck python scripts/sanitize_run.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from gknextengine_b200 import CudaPathTracingRenderer, GkcuFrameParams, host
from helpers import random_rays

for (W, H) in ((96, 56), (101, 57)):          # TMA stencils / plain-load stencils
    for name, sky_on in (("CornellBox.proc", False), ("RTIO.proc", True)):
        scene = host.Scene(name, samples=2, bounces=4, taa=True, denoiser=True)
        if sky_on: scene.env(has_sky=True, has_sun=True)
        scene.update_nodes()
        sky = host.make_sky(64, 32) if sky_on else None
        r = CudaPathTracingRenderer(0)
        r.CreateSwapChain((W, H))
        r.OnPostLoadScene(scene, sky)
        r.set_lights(scene)
        if sky_on: r.set_sky_sh(0, host.project_sh(sky))
        p = GkcuFrameParams(temporal_frames=16, spp_world=1)
        for f in range(1, 4):
            r.PreRender()
            r.Render(scene.ubo(W, H, f), p)
            r.BeforeNextFrame()
        r.Render(scene.ubo(W, H, 4), GkcuFrameParams(temporal_frames=16, spp_world=1, tracer=1))   # megakernel
        r.Render(scene.ubo(W, H, 5), GkcuFrameParams(temporal_frames=16, spp_world=2, spp_rank=0, skip_post=1))
        r.Resolve(scene.ubo(W, H, 5), GkcuFrameParams(temporal_frames=16, spp_world=2, spp_rank=0))
        r.BeforeNextFrame()
        r.trace(random_rays(2000, seed=3))
        r.trace(random_rays(2000, seed=4), any_hit=True)
        if name.startswith("Cornell") and W == 96:
            r.probes_voxelize()
            _, v, _ = r.probes_read()
            act = np.nonzero((v["distanceToSolid_gg_z01"] & 0xFF) < 8)[0]
            r.probes_bake(scene.ubo(W, H, 1), int(act.min()), 4096)
            r.probes_bake(scene.ubo(W, H, 1), int(act.min()), 4096)
            r.refit_model(0, scene.perturb_model(0, 1))
            r.PreRender(); r.Render(scene.ubo(W, H, 6), p); r.BeforeNextFrame()
        print(name, W, H, r.counters())
        r.close()
print("sanitize run finished")

# ---- round 2 additions: visibility front end, quantised nodes, tolerance stencil, unchanged-node skip, strip-parallel tail (one rank: its own peer)
for fmt in ("4", "8"):
    os.environ["GKCU_NODE_FORMAT"] = fmt
    W, H = 96, 56
    scene = host.Scene("RTIO.proc", samples=2, bounces=4, taa=True, denoiser=True)
    scene.update_nodes()
    r = CudaPathTracingRenderer(0)
    r.CreateSwapChain((W, H))
    r.OnPostLoadScene(scene, host.make_sky(64, 32))
    pv = GkcuFrameParams(temporal_frames=16, spp_world=1, primary_visibility=1)
    for f in range(1, 3):                       # the second frame culls against the first frame's depth
        r.Render(scene.ubo(W, H, f), pv)
        r.BeforeNextFrame()
    print("visibility front end", fmt, r.gpu_driven_stats())
    r.set_fast_stencil(True)
    r.set_skip_unchanged_nodes(True)
    r.set_kernel_timing(False)
    for f in range(3, 5):
        r.AfterUpdateScene(scene)
        r.PreRender()
        r.Render(scene.ubo(W, H, f), GkcuFrameParams(temporal_frames=16, spp_world=1))
        r.BeforeNextFrame()
    r.set_fast_stencil(False)
    for what in range(r.PEER_KINDS):
        r.peer_import(0, what, ptr=r.peer_export(what)[1])
    for f in range(5, 7):
        ps = GkcuFrameParams(temporal_frames=16, spp_world=1, spp_rank=0, skip_post=1)
        r.Render(scene.ubo(W, H, f), ps)
        r.ResolvePostStrip(scene.ubo(W, H, f), ps)
        r.BeforeNextFrame()
    print("node format", fmt, r.counters())
    r.close()

# ---- round 2, third session: TLAS refit / rebuild policy with moving instances, wavefront hand-over, segmented tail, variant knobs
os.environ["GKCU_NODE_FORMAT"] = "4"
os.environ["GKCU_TAIL_DIV"] = "4"          # the wavefront engages at this size and hands a wide tail over
os.environ["GKCU_TAIL_SEGMENTS"] = "2:3"
W, H = 224, 128
scene = host.Scene("LegoBricks.synth:600", samples=2, bounces=4, taa=True, denoiser=True)
scene.update_nodes()
r = CudaPathTracingRenderer(0)
r.CreateSwapChain((W, H))
r.OnPostLoadScene(scene, host.make_sky(64, 32))
p = GkcuFrameParams(temporal_frames=16, spp_world=1)
for f in range(1, 6):
    scene.jitter_instances(f, 0.05)
    scene.update_nodes()
    r.AfterUpdateScene(scene)
    if f == 4: r.set_tlas_refit(True, 2)
    r.PreRender()
    r.Render(scene.ubo(W, H, f), p)
    r.BeforeNextFrame()
r.set_tlas_refit(False, 0)
r.PreRender()
r.trace(random_rays(2000, seed=5, center=(0, 1.5, 0), spread=8.0))
print("tlas refit / segmented tail", r.tlas_stats(), r.counters())
r.close()
del os.environ["GKCU_TAIL_SEGMENTS"]
# the default tail: slots ordered by remaining samples (k_tail_hist / scan / scatter), then the fused vote-scheduled variant
for fused in ("0", "1"):
    os.environ["GKCU_TAIL_FUSED"] = fused
    r = CudaPathTracingRenderer(0)
    r.CreateSwapChain((W, H))
    r.OnPostLoadScene(scene, host.make_sky(64, 32))
    for f in range(1, 3):
        r.PreRender()
        r.Render(scene.ubo(W, H, f), p)
        r.BeforeNextFrame()
    print("ordered tail, fused =", fused, r.counters())
    r.close()
del os.environ["GKCU_TAIL_FUSED"], os.environ["GKCU_TAIL_DIV"]
