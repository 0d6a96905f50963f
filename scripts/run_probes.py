"""Probe-grid report: GPU voxelisation + bake times next to the CPU oracle, on both procedural scenes.
   python scripts/run_probes.py  ->  gpurun_out/probes_report.json"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_lib as O
from gknextengine_b200 import CudaPathTracingRenderer, host

def timed(fn, sync):
    sync(); t = time.perf_counter(); fn(); sync(); return (time.perf_counter() - t) * 1e3

report = {}
for name, sky_on, sun in (("CornellBox.proc", False, False), ("RTIO.proc", True, True)):
    scene = host.Scene(name, samples=1, bounces=4)
    if sky_on: scene.env(has_sky=True, has_sun=sun)
    scene.update_nodes()
    sky = host.make_sky(256, 128) if sky_on else None
    r = CudaPathTracingRenderer(0); r.CreateSwapChain((640, 360)); r.OnPostLoadScene(scene, sky); r.set_lights(scene)
    osc = O.OracleScene(scene, sky); osc.set_lights(scene)
    if sky_on:
        sh = host.project_sh(sky); r.set_sky_sh(0, sh); osc.set_sky_sh(0, sh)
    r.PreRender(); r.probes_voxelize(); r.BeforeNextFrame()  # warm-up
    vox_gpu = min(timed(r.probes_voxelize, r.BeforeNextFrame) for _ in range(3))
    t = time.perf_counter(); osc.probes_voxelize(); vox_cpu = (time.perf_counter() - t) * 1e3
    _, gv, _ = r.probes_read(); _, ov, _ = osc.probes_read()
    d0 = gv["distanceToSolid_gg_z01"] & 0xFF
    active = np.nonzero(d0 < 8)[0]
    ubo = scene.ubo(640, 360, 1)
    per_frame = (192 * 192 * 48 // 64 // 120) * 64
    # one full sweep = 120 dispatches of the reference's per-frame size
    sweep = lambda: [r.probes_bake(ubo, f * per_frame, per_frame) for f in range(120)]
    sweep(); r.BeforeNextFrame()
    bake_gpu = min(timed(sweep, r.BeforeNextFrame) for _ in range(2))
    t = time.perf_counter()
    for f in range(120): osc.probes_bake(ubo, f * per_frame, per_frame)
    bake_cpu = (time.perf_counter() - t) * 1e3
    for f in range(120): osc.probes_bake(ubo, f * per_frame, per_frame)
    for f in range(120): osc.probes_bake(ubo, f * per_frame, per_frame)  # oracle has had 3 sweeps like the GPU (warm-up + 2 timed)
    gc, gv2, _ = r.probes_read(); oc, ov2, _ = osc.probes_read()
    rays = len(active) * 96
    report[name] = dict(voxels=int(len(gv)), surface_probes=int(len(active)), voxel_parity=bool((gv == ov).all()), cube_parity=bool((gc == oc).all()),
                        voxelize_gpu_ms=vox_gpu, voxelize_cpu_ms=vox_cpu, bake_sweep_gpu_ms=bake_gpu, bake_per_frame_gpu_ms=bake_gpu / 120,
                        bake_sweep_cpu_ms=bake_cpu, bake_rays_per_sweep=int(rays), bake_gpu_mrays_s=rays / bake_gpu / 1e3,
                        cpu_threads=os.cpu_count())
    r.close()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(report, open(os.path.join(ROOT, "gpurun_out", "probes_report.json"), "w"), indent=1)
print(json.dumps(report, indent=1))
