"""Bring-up diagnostics on a GPU box: CUDA renderer vs the CPU oracle on small configs (prints, does not assert)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O  # noqa: E402
from gknextengine_b200 import RT, CudaPathTracingRenderer, GkcuFrameParams, host  # noqa: E402
from gknextengine_b200._abi import RAY_DTYPE  # noqa: E402


def compare(name, a, b):
    a32, b32 = a.astype(np.float64), b.astype(np.float64)
    neq = (a != b)
    if neq.ndim == 3:
        neq = neq.any(axis=2)
    print(f"  {name:18s} mismatching pixels {neq.mean():.3e}  max|diff| {np.nanmax(np.abs(a32 - b32)):.4g}")


def run(scene_name, W, H, samples, bounces, tracer, sky=None, frames=1, denoiser=True, env=None):
    print(f"== {scene_name} {W}x{H} spp={samples} bounces={bounces} tracer={tracer}")
    scene = host.Scene(scene_name, samples=samples, bounces=bounces, taa=False, denoiser=denoiser)
    if env:
        scene.env(*env)
    scene.update_nodes()
    print("  ", scene.counts())
    r = CudaPathTracingRenderer(0)
    r.CreateSwapChain((W, H))
    t0 = time.time()
    r.OnPostLoadScene(scene, sky)
    r.PreRender()
    print(f"   AS build {1e3 * (time.time() - t0):.1f} ms wall, timings {r.timings()}")
    osc = O.OracleScene(scene, sky)
    ofr = O.OracleFrame(W, H)
    p = GkcuFrameParams(temporal_frames=16, spp_world=1, tracer=tracer)
    for f in range(1, frames + 1):
        ubo = scene.ubo(W, H, f)
        r.Render(ubo, p)
        r.BeforeNextFrame()
        t0 = time.time()
        ofr.render(osc, ubo, p)
        print(f"   frame {f}: gpu {r.timings()} counters {r.counters()} | oracle {time.time() - t0:.2f}s {ofr.counters()}")
        for k in ("OBJEDCTID_0", "EXT_PRIMITIVEID", "ALBEDO", "NORMAL", "MOTIONVECTOR", "SINGLE_DIFFUSE", "SINGLE_SPECULAR",
                  "ACCUMLATE_DIFFUSE", "DENOISED"):
            compare(k, r.readback(RT[k]), ofr.image(RT[k]))
    # random rays through the intersector
    rng = np.random.default_rng(1)
    n = 20000
    rays = np.zeros(n, dtype=RAY_DTYPE)
    o = rng.uniform(-3, 3, (n, 3)).astype(np.float32) + np.array([0, 3, 0], np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays["ox"], rays["oy"], rays["oz"] = o.T
    rays["dx"], rays["dy"], rays["dz"] = d.T
    rays["tmin"] = 1e-3
    rays["tmax"] = 1000
    hg, hc = r.trace(rays), osc.trace(rays)
    print(f"   trace_rays: inst mismatch {(hg['inst'] != hc['inst']).mean():.2e} prim mismatch {(hg['prim'] != hc['prim']).mean():.2e} "
          f"t bit-exact {(hg['t'] == hc['t']).mean():.4f} hits {(hc['inst'] != 0xFFFFFFFF).mean():.3f}")
    ag, ac = r.trace(rays, any_hit=True), osc.trace(rays, any_hit=True)
    print(f"   any-hit mismatch {(ag['inst'] != ac['inst']).mean():.2e}")
    r.close()


if __name__ == "__main__":
    sky = host.make_sky(512, 256)
    for builder in ("lbvh", "sah"):
        os.environ["GKCU_BLAS_BUILDER"] = builder
        print("#### BLAS builder", builder)
        run("CornellBox.proc", 320, 180, 1, 4, 1)
        run("CornellBox.proc", 320, 180, 2, 4, 0, frames=2)
        run("RTIO.proc", 320, 180, 2, 4, 0, sky=sky)
    run("RTIO.proc", 320, 180, 2, 4, 1, sky=sky, env=(True, True))
