"""The bench contract on CPU: the reference arm of bench.py runs here (it times the CPU port), prints exactly one JSON line with the keys the
driver reads, and the committed round profile of the CUDA arm carries the same contract plus roofline / cpu_baseline / e2e / clocks."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "e2e"}


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"], capture_output=True, text=True,
                         timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip().startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert BASE_KEYS <= set(j) and j["impl"] == "reference" and j["metric"] and j["unit"] == "Mrays/s" and j["higher_is_better"] is True
    assert j["value"] > 0 and j["n_gpus"] == 1 and j["steps"] == 1 and "workload" in j["config"] and j["vs_baseline"] is None
    cb = j["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == j["value"] and cb["sample"]
    assert j["e2e"] == {"value": j["value"], "unit": j["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_committed_round_profile_carries_the_contract():
    path = os.path.join(ROOT, "profiles", "r01_bench_n1.json")
    j = json.loads([ln for ln in open(path).read().splitlines() if ln.strip().startswith("{")][-1])
    assert BASE_KEYS <= set(j) and j["n_gpus"] == 1 and j["warmup"] >= 3 and j["gpu_launches"] > 0 and j["data"].startswith("synthetic")
    r = j["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["traffic"]
    e = j["e2e"]
    assert 0 < e["value"] <= j["value"] * 1.001 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    c = j["clocks"]
    assert c["sm_mhz"] > 0 and c["sm_max_mhz"] >= c["sm_mhz"] and not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    cb = j["cpu_baseline"]
    assert cb["value"] > 0 and cb["cores"] >= 1 and cb["kind"] in ("port", "reference") and cb["sample"]
    for k in ("reproject", "denoise"):
        assert j["roofline_stencils"][k]["bound"] == "issue" and 0 < j["roofline_stencils"][k]["issue_frac"] < 1


def test_committed_round2_profile_carries_the_contract():
    """profiles/r02_bench_n1.json: the line of the final round-2 code. The end-to-end leg runs the library defaults (no per-launch event brackets), so it
    may read above the device-resident value, whose timed region carries them for the roofline; both must be consistent with the frame without
    brackets. The multi-GPU lines of the same code carry the strong-scaling block."""
    load = lambda name: json.loads([ln for ln in open(os.path.join(ROOT, "profiles", name)).read().splitlines() if ln.strip().startswith("{")][-1])
    j = load("r02_bench_n1.json")
    assert BASE_KEYS <= set(j) and j["n_gpus"] == 1 and j["warmup"] >= 3 and j["gpu_launches"] > 0 and j["data"].startswith("synthetic")
    assert j["metric"] == "Mrays/s" and j["dtype"] == "f32" and j["vs_baseline"] is None and "workload" in j["config"] and "l2" in j["config"]
    r = j["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["traffic"] > 0 and r["kernel"] == "k_trace"
    assert 0.5 < r["share_of_step"] < 0.75 and r["launches"] == 42 * j["steps"]
    e, plain = j["e2e"], j["without_kernel_events"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] == 1920 * 1080 * 8 and "defaults" in e["note"]
    assert plain["ms_per_step"] < j["ms_per_step"] and plain["ms_per_step"] <= e["ms_per_step"] < 1.1 * j["ms_per_step"]
    c = j["clocks"]
    assert c["samples"] >= 3 and c["sm_mhz"] > 0 and c["sm_max_mhz"] >= c["sm_mhz"] and not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    cb = j["cpu_baseline"]
    assert cb["value"] > 0 and cb["cores"] >= 1 and cb["kind"] == "reference" and "tinybvh" in cb["sample"]
    assert j["tlas_update"]["refits"] > j["tlas_update"]["builds"] >= 1
    s1 = j["strong"]
    assert s1["scaling"] == "strong" and s1["width"] == 3840 and s1["spp_per_gpu"] == 64
    for n in (2, 8):
        m = load(f"r02_bench_n{n}.json")
        assert m["n_gpus"] == n and m["scaling"] == "weak" and 0.97 * n < m["value"] / j["value"] < 1.03 * n
        sn = m["strong"]
        assert sn["scaling"] == "strong" and sn["spp_per_gpu"] == 64 // n and 0.85 * n < s1["ms_per_step"] / sn["ms_per_step"] <= n


def test_cpu_baseline_reference_and_port_trace_the_same_rays_and_threads_do_not_change_results():
    """bench.py's CPU baseline: the compiled reference shaders (kind "reference", multi-threaded over rows) and the oracle port must count the
    same rays on the same band of the bench frame, and the thread count must not change a single pixel."""
    import numpy as np
    sys.path.insert(0, ROOT)
    import bench
    from gknextengine_b200 import GkcuFrameParams, host
    scene = host.Scene(bench.SCENE, samples=2, bounces=bench.BOUNCES, taa=True, temporal_frames=16, denoiser=True)
    scene.update_nodes()
    sky = host.make_sky(64, 32)
    try:
        ref = bench.CompiledReference(scene, sky, query="native")
    except FileNotFoundError:
        import pytest
        pytest.skip("oracle/_ref not built (reference not mounted)")
    ubo = scene.ubo(bench.WIDTH, bench.HEIGHT, 3)
    rays1, _ = ref.sample(ubo, 1, 2, threads=1)
    band = slice(bench.HEIGHT // 2 - 1, bench.HEIGHT // 2 + 1)
    first = {k: v[band].copy() for k, v in ref.images.items()}
    rays4, _ = ref.sample(ubo, 1, 2, threads=4)
    assert rays1 == rays4 and all(np.array_equal(first[k].view(np.uint8), ref.images[k][band].view(np.uint8)) for k in first)
    prays, _ = bench.cpu_port_sample(scene, sky, ubo, GkcuFrameParams(temporal_frames=16, spp_world=1), 1, 2)
    assert prays == rays1 and rays1 > 2 * bench.WIDTH * 3
    assert first["SINGLE_DIFFUSE"].astype(np.float32).max() > 0


def test_cpu_baseline_runs_on_the_references_own_tinybvh_and_scales_over_pixel_groups():
    """The timed CPU arm answers ray queries with the reference's vendored tinybvh (src/Assets/CPUAccelerationStructure.cpp:172-284 usage) and hands
    out 8 x 8 pixel groups, so a thin band still occupies every core; thread count must not change a pixel, and the ray count must agree with the
    native-intersector build to within the edge-tie budget (the two intersectors differ only on triangle edges)."""
    import numpy as np
    sys.path.insert(0, ROOT)
    import bench
    from gknextengine_b200 import host
    scene = host.Scene(bench.SCENE, samples=2, bounces=bench.BOUNCES, taa=True, temporal_frames=16, denoiser=True)
    scene.update_nodes()
    sky = host.make_sky(64, 32)
    try:
        ref = bench.CompiledReference(scene, sky)
    except FileNotFoundError:
        import pytest
        pytest.skip("oracle/_ref not built (reference not mounted)")
    if "tinybvh" not in ref.query:
        import pytest
        pytest.skip("tinybvh build of the compiled reference not present")
    ubo = scene.ubo(bench.WIDTH, bench.HEIGHT, 3)
    rays1, _ = ref.sample(ubo, 1, 8, threads=1)
    band = slice(bench.HEIGHT // 2 - 4, bench.HEIGHT // 2 + 4)
    first = {k: v[band].copy() for k, v in ref.images.items()}
    rays8, _ = ref.sample(ubo, 1, 8, threads=8)
    assert rays1 == rays8 and all(np.array_equal(first[k].view(np.uint8), ref.images[k][band].view(np.uint8)) for k in first)
    native = bench.CompiledReference(scene, sky, query="native")
    nrays, _ = native.sample(ubo, 1, 8, threads=8)
    assert abs(nrays - rays1) <= 2e-3 * nrays
