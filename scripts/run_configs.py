"""Runs BASELINE.json configs 1, 3, 4 (and the single-GPU leg of 5) on one B200 and writes the numbers SURVEY 8(d) asks for
to gpurun_out/configs_report.json / .md. Config 2 and the multi-GPU legs are bench.py's job. The oracle is used as the checker."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O  # noqa: E402
from gknextengine_b200 import RT, CudaPathTracingRenderer, GkcuFrameParams, host  # noqa: E402
from helpers import half_ulp_diff, mismatch_fraction, rel_rmse  # noqa: E402

PEAK = 6548.8
report = {}
quick = "--quick" in sys.argv


def rays(c):
    return c["primary_rays"] + c["closest_hit_rays"] + c["any_hit_rays"]


def config1():
    W, H, N = 640, 360, 64 if quick else 256
    scene = host.Scene("CornellBox.proc", samples=1, bounces=4, taa=False)
    scene.update_nodes()
    r = CudaPathTracingRenderer(0)
    r.set_kernel_timing(True)  # per-launch event brackets: kernel_stats() needs them
    r.CreateSwapChain((W, H))
    r.OnPostLoadScene(scene)
    osc, ofr = O.OracleScene(scene), O.OracleFrame(W, H)
    p = GkcuFrameParams(temporal_frames=16, spp_world=1)
    acc_g, acc_c = np.zeros((H, W, 3)), np.zeros((H, W, 3))
    id_mis, prim_mis, exact, gpu_ms, cpu_s, nrays = 0.0, 0.0, 0, [], 0.0, 0
    for f in range(1, N + 1):
        ubo = scene.ubo(W, H, f)
        r.Render(ubo, p)
        r.BeforeNextFrame()
        t0 = time.perf_counter()
        ofr.render(osc, ubo, p, run_post=False)
        cpu_s += time.perf_counter() - t0
        g, c = r.readback(RT["SINGLE_DIFFUSE"]), ofr.image(RT["SINGLE_DIFFUSE"])
        exact += int(np.array_equal(g.view(np.uint16), c.view(np.uint16)))
        acc_g += g.astype(np.float64)[..., :3]
        acc_c += c.astype(np.float64)[..., :3]
        id_mis = max(id_mis, mismatch_fraction(r.readback(RT["OBJEDCTID_0"]), ofr.image(RT["OBJEDCTID_0"])))
        prim_mis = max(prim_mis, mismatch_fraction(r.readback(RT["EXT_PRIMITIVEID"]), ofr.image(RT["EXT_PRIMITIVEID"])))
        gpu_ms.append(r.timings()["frame_ms"])
        nrays += rays(r.counters())
    report["config1_cornell_640x360_1spp_4b"] = {
        "frames": N, "primary_id_mismatch_fraction_max": id_mis, "primitive_id_mismatch_fraction_max": prim_mis,
        "frames_with_bit_identical_SINGLE_DIFFUSE": exact, "radiance_rel_rmse_after_N_frames": rel_rmse(acc_g, acc_c),
        "gpu_ms_per_frame": float(np.mean(gpu_ms)), "gpu_mrays_s": nrays / (sum(gpu_ms) * 1e3),
        "cpu_oracle_ms_per_frame": 1e3 * cpu_s / N, "cpu_cores": os.cpu_count()}
    r.close()


def config3():
    frames = 60 if quick else 300
    sky = host.make_sky(1024, 512)

    def run(W, H, check):
        scene = host.Scene("RTIO.proc", samples=1, bounces=5, taa=True, temporal_frames=16, denoiser=True)
        scene.update_nodes()
        r = CudaPathTracingRenderer(0)
        r.CreateSwapChain((W, H))
        r.OnPostLoadScene(scene, sky)
        p = GkcuFrameParams(temporal_frames=16, spp_world=1)
        osc = ofr = None
        if check:
            osc, ofr = O.OracleScene(scene, sky), O.OracleFrame(W, H)
        times, passes, probes = [], [], {}
        for f in range(1, frames + 1):
            ubo = scene.ubo(W, H, f, model_rotation_deg=(f / 60.0) * 20.0)  # gkNextMotionBenchmark.cpp:50-51
            r.PreRender()
            r.Render(ubo, p)
            r.BeforeNextFrame()
            t = r.timings()
            times.append(t["frame_ms"] + t["tlas_update_ms"])
            passes.append(t)
            if check:
                ofr.render(osc, ubo, p)
                if f in (1, 16, frames // 2, frames):
                    g, c = r.readback(RT["DENOISED"]), ofr.image(RT["DENOISED"])
                    ok = np.isfinite(g.astype(np.float32)).all(axis=2) & np.isfinite(c.astype(np.float32)).all(axis=2)
                    probes[f] = {"rel_rmse": rel_rmse(g[ok], c[ok]), "pixels_not_bit_identical": float(((g.view(np.uint16) != c.view(np.uint16)).any(axis=2) & ok).mean())}
        r.close()
        return times, passes, probes

    times, passes, _ = run(1920, 1080, False)
    _, _, probes = run(480, 270, True)
    px = 1920 * 1080
    rp = float(np.mean([p["reproject_pass_ms"] for p in passes[5:]]))
    dn = float(np.mean([p["compose_pass_ms"] for p in passes[5:]]))
    report["config3_temporal_denoise_1080p_1spp_orbit"] = {
        "scene": "RTIO.proc stand-in (luxball.glb is absent from the mount)", "frames": frames,
        "ms_per_frame_mean": float(np.mean(times[5:])), "ms_per_frame_p95": float(np.percentile(times[5:], 95)),
        "rt_pass_ms": float(np.mean([p["rt_pass_ms"] for p in passes[5:]])), "tlas_update_ms": float(np.mean([p["tlas_update_ms"] for p in passes[5:]])),
        "reproject_pass_ms": rp, "compose_pass_ms": dn,
        "reproject_GBps_at_98B_per_pixel": 98 * px / (rp * 1e-3) / 1e9, "reproject_frac_of_hbm_peak": 98 * px / (rp * 1e-3) / 1e9 / PEAK,
        "denoise_GBps_at_44B_per_pixel": 44 * px / (dn * 1e-3) / 1e9, "denoise_frac_of_hbm_peak": 44 * px / (dn * 1e-3) / 1e9 / PEAK,
        "parity_480x270_RT_DENOISED_vs_oracle": probes}


def config4():
    n_inst = 20000 if quick else 100000
    W, H = 1920, 1080
    scene = host.Scene(f"LegoBricks.synth:{n_inst}", samples=1, bounces=4, taa=True, denoiser=True)
    scene.update_nodes()
    sky = host.make_sky(1024, 512)
    r = CudaPathTracingRenderer(0)
    r.CreateSwapChain((W, H))
    t0 = time.perf_counter()
    r.OnPostLoadScene(scene, sky)
    r.PreRender()
    load_ms = 1e3 * (time.perf_counter() - t0)
    p = GkcuFrameParams(temporal_frames=16, spp_world=1)
    osc = O.OracleScene(scene, sky)
    rows = []
    for f in range(1, 13):
        scene.jitter_instances(f, 0.01)
        scene.update_nodes()
        t0 = time.perf_counter()
        if f % 2:
            r.AfterUpdateScene(scene)         # every NodeProxy (what the reference does, Scene.cpp:443-492)
        else:
            r.AfterUpdateSceneSparse(scene)   # only the proxies that changed
        upload_ms = 1e3 * (time.perf_counter() - t0)
        model = 1 + (f % 16)
        sv = scene.perturb_model(model, f)
        r.refit_model(model, sv)
        osc.refit_model(model, sv)
        r.PreRender()
        r.Render(scene.ubo(W, H, f), p)
        r.BeforeNextFrame()
        t, c, k = r.timings(), r.counters(), r.kernel_stats()
        rows.append(dict(sparse=(f % 2 == 0), changed=scene.changed()[1], tlas=t["tlas_update_ms"], refit=t["blas_refit_ms"], rt=t["rt_pass_ms"], frame=t["frame_ms"], upload=upload_ms, rays=rays(c),
                         trace=k["trace_ms"] + k["tail_ms"] + k["primary_ms"]))
    rows = rows[2:]
    cnt = scene.counts()
    # parity of the dynamic AS on a ray batch against the oracle (rebuilt from scratch)
    osc.update_nodes(scene)
    from helpers import random_rays
    rr = random_rays(100000, seed=5, center=(0, 2, 0), spread=20.0)
    rr["dy"] = -np.abs(rr["dy"])
    hg, hc = r.trace(rr), osc.trace(rr)
    report["config4_dynamic_as_lego_1080p"] = {
        "instances": cnt["nodes"], "instanced_triangles": cnt["instanced_triangles"], "models": cnt["models"],
        "scene_load_blas_build_tlas_ms_wall": load_ms, "blas_build_ms": r.timings()["blas_build_ms"],
        "tlas_rebuild_ms": float(np.mean([x["tlas"] for x in rows])), "blas_refit_ms": float(np.mean([x["refit"] for x in rows])),
        "node_upload_full_ms_wall": float(np.mean([x["upload"] for x in rows if not x["sparse"]])),
        "node_upload_sparse_ms_wall": float(np.mean([x["upload"] for x in rows if x["sparse"]])),
        "proxies_changed_per_frame": float(np.mean([x["changed"] for x in rows])),
        "rt_pass_ms": float(np.mean([x["rt"] for x in rows])), "trace_kernels_ms": float(np.mean([x["trace"] for x in rows])),
        "frame_ms": float(np.mean([x["frame"] for x in rows])),
        "mrays_s_rt_pass": float(np.mean([x["rays"] / (x["rt"] * 1e3) for x in rows])),
        "hit_parity_vs_oracle_100k_rays": {"instance_mismatch": float((hg["inst"] != hc["inst"]).mean()), "t_bit_exact": float((hg["t"].view(np.uint32) == hc["t"].view(np.uint32)).mean())}}
    r.close()


def config5_single():
    W, H, spp = 3840, 2160, 16 if quick else 64
    scene = host.Scene("RTIO.proc", samples=spp, bounces=8, taa=True, denoiser=True)
    scene.update_nodes()
    r = CudaPathTracingRenderer(0)
    r.CreateSwapChain((W, H))
    r.OnPostLoadScene(scene, host.make_sky(1024, 512))
    p = GkcuFrameParams(temporal_frames=16, spp_world=1)
    ms, nr = [], 0
    for f in range(1, 4):
        r.PreRender()
        r.Render(scene.ubo(W, H, f), p)
        r.BeforeNextFrame()
        if f > 1:
            ms.append(r.timings()["frame_ms"])
            nr += rays(r.counters())
    report["config5_4k_64spp_single_gpu_leg"] = {"scene": "RTIO.proc stand-in (qx50.glb is absent from the mount)", "spp": spp,
                                                 "ms_per_frame": float(np.mean(ms)), "mrays_s": nr / (sum(ms) * 1e3)}
    r.close()


if __name__ == "__main__":
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    only = [a for a in sys.argv[1:] if a.startswith("config")]
    for fn in (config1, config3, config4, config5_single):
        if only and fn.__name__ not in only:
            continue
        t0 = time.time()
        fn()
        print(fn.__name__, f"{time.time() - t0:.1f}s", flush=True)
    with open(os.path.join(ROOT, "gpurun_out", "configs_report.json"), "w") as f:
        json.dump(report, f, indent=1)
    print(json.dumps(report, indent=1))
