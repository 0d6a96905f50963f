#!/usr/bin/env python
"""bench.py — Mrays/s and ms/frame of the path-tracing frame on N B200s (BASELINE.json metric), one JSON line.

Workload (config.workload): BASELINE.json configs[1] — the RayTracingInOneWeekend procedural scene (RTIO.proc,
src/Runtime/SceneList.cpp:117-178), 1920x1080, 16 spp, 8 bounces, sky, steady state (no node motion). A "step" is one
frame through PathTracingRenderer::Render's pass sequence: TLAS update -> rt pass -> reproject pass -> compose pass.
With N > 1 every rank holds a scene replica and traces 16 spp of a 16*N-spp frame (weak scaling over samples per pixel,
SURVEY 8e); the un-normalised radiance sums are reduced to rank 0 with NCCL and rank 0 resolves + runs the post passes.

  value          rays/s of the whole job, inputs resident in HBM, CUDA-event timed on the renderer's stream, max over ranks
  e2e            the same through the public API with HOST buffers: NodeProxy[] + UBO H2D from pinned memory and the
                 RT_DENOISED D2H read every step, wall-clock between synchronisations
  roofline       dominant kernel (k_trace) against the measured HBM peak, see DESIGN.md for the byte accounting
  cpu_baseline   the reference's RT-pass shaders compiled as C++ (oracle/_ref, kind "reference") or, without them, the CPU oracle port (kind "port")
                 on all host cores, bounded sample of the same frame

--impl reference times that CPU baseline alone (the reference's Slang/Vulkan path itself cannot run without slangc/Vulkan; its shader text
compiled as C++ can).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WIDTH, HEIGHT, SPP, BOUNCES = 1920, 1080, 16, 8
SKY_PATH = None  # --sky: a Radiance .hdr equirect (e.g. the reference's assets/textures/river_road_2.hdr) instead of the procedural stand-in
SCENE = "RTIO.proc"
# algorithmic bytes (DESIGN.md "Byte accounting"): per closest-hit ray 436 B, per any-hit ray 100 B for the
# trace+shade pair; the share k_trace itself moves is the ray record it reads and the hit record it writes.
TRACE_BYTES_PER_RAY = 32 + 4 + 16 + 4
PAIR_BYTES_CLOSEST, PAIR_BYTES_ANY = 436, 100
PRIMARY_BYTES_PER_PIXEL, REPROJECT_BYTES_PER_PIXEL, DENOISE_BYTES_PER_PIXEL = 36 + 2, 98, 44
# Both stencil kernels are bound by instruction issue, not by HBM (ncu: profiles/r02_ncu_table.md); the bench reports their algorithmic GB/s only.


def bench_sky(host):
    """Sky id 0 of the workload (Engine.cpp:855): the file given with --sky, else the procedural stand-in."""
    if SKY_PATH:
        return host.load_hdr(SKY_PATH)
    return host.make_sky(1024, 512)


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons with nvidia-smi while the timed region runs."""

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu_index = gpu_index
        self.samples = []
        self.proc = None

    def run_nvml(self):
        """NVML polled every 10 ms: the timed region of the default run is ~0.3 s, nvidia-smi's loop would see it once or twice."""
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(self.gpu_index)
        mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        bits = [(getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8), 2), (getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40), 3),
                (getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20), 4), (getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4), 5)]
        self.nvml_running = True
        while self.nvml_running:
            r = get_reasons(h)
            row = [str(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), str(mx), "", "", "", ""]
            for mask, col in bits:
                row[col] = "Active" if (r & mask) else "Not Active"
            self.samples.append(row)
            time.sleep(0.01)

    def run(self):
        try:
            self.run_nvml()
            return
        except Exception:
            pass
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu_index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                parts = [p.strip() for p in line.split(",")]
                if len(parts) >= 6:
                    self.samples.append(parts)
        except Exception:
            pass

    def stop(self):
        self.nvml_running = False
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                mx = max(mx, float(s[1]))
            except ValueError:
                continue
            for n, v in zip(names, s[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def cpu_port_sample(scene, sky, ubo, params, bands, band_rows, threads=0):
    """Times the CPU oracle (reference-semantics port) on `bands` evenly spaced bands of `band_rows` rows of the frame."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O

    osc = O.OracleScene(scene, sky)
    ofr = O.OracleFrame(WIDTH, HEIGHT)
    rays, secs = 0, 0.0
    for b in range(bands):
        y0 = int((b + 0.5) * HEIGHT / bands) - band_rows // 2
        ofr.render(osc, ubo, params, threads=threads, rows=(y0, y0 + band_rows), run_post=False)
        c = ofr.counters()
        rays += c["primary"] + c["closest"] + c["any"]
        secs += ofr.rt_seconds()
    return rays, secs


class CompiledReference:
    """The reference's own RT pass - Util.BufferClear + Core.PathTracing with the Shading.slang / AmbientCube.slang classes they run - compiled
    as C++ from the reference's sources into oracle/_ref/libref_pathtrace.so (oracle/ref_pathtrace.cpp; built when the reference is mounted, the
    binary travels with the repo). Ray queries, the texture sampler and the visibility buffer of the raster pass are served by the CPU oracle's
    intersector (they are hardware in the reference): the reference's vendored tinybvh when oracle/_ref holds that build. Multi-threaded over
    8 x 8 pixel groups."""

    def __init__(self, scene, sky, query="tinybvh"):
        import ctypes as C
        import numpy as np
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib as O
        from gknextengine_b200 import RT
        from gknextengine_b200._abi import GK_CUBE_COUNT
        from gknextengine_b200.renderer import CudaPathTracingRenderer as R
        # the reference's own CPU ray query (vendored tinybvh, used like src/Assets/CPUAccelerationStructure.cpp:172-284) when that build exists
        path = os.path.join(os.path.dirname(O.TINYBVH), "libref_pathtrace_tinybvh.so")
        oracle_path, self.query = O.TINYBVH, "the reference's vendored tinybvh 1.3.8"
        if query != "tinybvh" or not (os.path.exists(path) and os.path.exists(O.TINYBVH)):
            path = os.path.join(os.path.dirname(O.TINYBVH), "libref_pathtrace.so")
            oracle_path, self.query = O.NATIVE, "the CPU oracle's BVH2"
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.osc = O.OracleScene(scene, sky, lib_path=oracle_path)  # first: the shader library links against this intersector library
        self.C, self.lib = C, C.CDLL(path)
        self.lib.ref_pathtrace_rows.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint64)]
        # zero probe state; the shader dereferences the probe arrays, so they must exist
        self.osc.set_probes(np.zeros(GK_CUBE_COUNT, R.CUBE_DTYPE), np.zeros(GK_CUBE_COUNT, R.VOXEL_DTYPE))
        self.osc.L.gko_cast_visibility_rows.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p]
        W, H = WIDTH, HEIGHT
        self.images = {"SINGLE_DIFFUSE": np.zeros((H, W, 4), np.float16), "SINGLE_SPECULAR": np.zeros((H, W, 4), np.float16), "ALBEDO": np.zeros((H, W, 4), np.float16),
                       "NORMAL": np.zeros((H, W, 4), np.float16), "OBJEDCTID_0": np.zeros((H, W), np.uint32), "MINIGBUFFER": np.zeros((H, W), np.uint32),
                       "MOTIONVECTOR": np.zeros((H, W, 2), np.float32), "PREV_DEPTHBUFFER": np.zeros((H, W), np.float32), "MOTIONMOMENT": np.zeros((H, W), np.uint16)}
        self.table = (C.c_void_p * 64)()
        for k, v in self.images.items():
            self.table[RT[k]] = v.ctypes.data

    def sample(self, ubo, bands, band_rows, threads=0):
        """(rays, seconds) of the RT pass on `bands` evenly spaced bands of `band_rows` rows; the visibility cast stands in for the raster pass
        and is timed and counted as the primary rays, like in the CUDA renderer's rt pass."""
        C = self.C
        counters = (C.c_uint64 * 2)()
        rays, secs = 0, 0.0
        for b in range(bands):
            y0 = int((b + 0.5) * HEIGHT / bands) - band_rows // 2
            t0 = time.perf_counter()
            self.osc.L.gko_cast_visibility_rows(self.osc._h, C.byref(ubo), WIDTH, HEIGHT, y0, y0 + band_rows, threads, self.images["MINIGBUFFER"].ctypes.data)
            self.lib.ref_pathtrace_rows(self.osc._h, C.byref(ubo), self.table, WIDTH, HEIGHT, y0, y0 + band_rows, threads, counters)
            secs += time.perf_counter() - t0
            rays += band_rows * WIDTH + counters[0] + counters[1]
        return rays, secs


def cpu_sample(scene, sky, ubo, params, bands, band_rows):
    """The CPU baseline of one frame sample: the compiled reference shaders when oracle/_ref holds them (kind "reference"), else the oracle
    port (kind "port"). Returns (rays, seconds, kind, description). The compiled-reference leg runs in a child process (no GPU, no torch)
    so that whatever happens there the port can still be timed and the bench line printed."""
    try:
        cmd = [sys.executable, os.path.abspath(__file__), "--cpu-leg", f"{bands},{band_rows}", "--workload", "config5" if WIDTH == 3840 else "config2"]
        if SKY_PATH:
            cmd += ["--sky", SKY_PATH]
        res = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
        j = json.loads([ln for ln in res.stdout.splitlines() if ln.startswith("{")][-1])
        return j["rays"], j["secs"], "reference", f"the reference's Core.PathTracing shader compiled as C++ (oracle/_ref/libref_pathtrace*.so), ray queries on {j['query']}"
    except Exception:  # library not built (reference not mounted when the repo was built) or unusable here: the port always exists
        rays, secs = cpu_port_sample(scene, sky, ubo, params, bands, band_rows)
        return rays, secs, "port", "CPU port of the reference shaders (oracle/)"


def cpu_leg(spec):
    """Child process of cpu_sample: frame 1 of the bench workload through the compiled reference shaders; prints {"rays", "secs"}."""
    from gknextengine_b200 import host
    bands, band_rows = (int(x) for x in spec.split(","))
    scene = host.Scene(SCENE, samples=SPP, bounces=BOUNCES, taa=True, temporal_frames=16, denoiser=True)
    scene.update_nodes()
    ref = CompiledReference(scene, bench_sky(host))
    ubo = scene.ubo(WIDTH, HEIGHT, 1)
    ref.sample(ubo, 2, 2)  # untimed: first touch of the image pages, thread start-up
    rays, secs = ref.sample(ubo, bands, band_rows)
    print(json.dumps({"rays": int(rays), "secs": secs, "query": ref.query}))


def reference_arm(args):
    """--impl reference: the reference's RT pass on all host cores (compiled shader text if present, else the CPU port); rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from gknextengine_b200 import GkcuFrameParams, host

    scene = host.Scene(SCENE, samples=SPP * args.gpus, bounces=BOUNCES, taa=True, temporal_frames=16, denoiser=True)
    scene.update_nodes()
    sky = bench_sky(host)
    params = GkcuFrameParams(temporal_frames=16, spp_world=1)
    cores = os.cpu_count() or 1
    bands, band_rows = 8, 16  # 128 of 1080 rows per step; work items are 8 x 8 pixel groups: 480 per band, enough for any host
    total_rays, total_s = 0, 0.0
    try:
        ref = CompiledReference(scene, sky)
        kind, what = "reference", f"the reference's Core.PathTracing shader compiled as C++ (oracle/_ref/libref_pathtrace*.so), ray queries on {ref.query}"
    except Exception:
        ref, kind, what = None, "port", "CPU port of the reference shaders (oracle/)"
    for step in range(args.warmup + args.steps):
        ubo = scene.ubo(WIDTH, HEIGHT, step + 1)
        rays, secs = ref.sample(ubo, bands, band_rows) if ref else cpu_port_sample(scene, sky, ubo, params, bands, band_rows)
        if step >= args.warmup:
            total_rays += rays
            total_s += secs
    value = total_rays / total_s / 1e6
    sample = f"{bands} bands x {band_rows} rows of the {WIDTH}x{HEIGHT} frame per step, rt pass only; {what}"
    print(json.dumps({
        "impl": "reference", "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total_s / max(1, args.steps), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(args.gpus, args.probes),
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def zero_copy_leg(r, scene, frame, sync_all, steps, tf, pinned_nodes, node_bytes):
    """The e2e loop with RT_DENOISED bound to imported external memory (gkcu_bind_external_image): what the in-engine adaptor does with the
    engine's exportable render image (src/Vulkan/RenderImage.hpp:30-37), so the 16.6 MB device->host read of the headless path disappears. Without
    Vulkan on the box the exporter is a cuMemCreate allocation shared through a POSIX descriptor. Returns the e2e_zero_copy block (None when
    cuda-python is missing)."""
    try:
        from cuda.bindings import driver as cu
    except Exception:
        return None
    from gknextengine_b200 import RT

    def ck(res):
        if res[0] != cu.CUresult.CUDA_SUCCESS:
            raise RuntimeError(str(res[0]))
        return res[1] if len(res) == 2 else res[1:]

    try:
        ck(cu.cuInit(0))
        prop = cu.CUmemAllocationProp()
        prop.type = cu.CUmemAllocationType.CU_MEM_ALLOCATION_TYPE_PINNED
        prop.location.type = cu.CUmemLocationType.CU_MEM_LOCATION_TYPE_DEVICE
        prop.location.id = int(os.environ.get("LOCAL_RANK", "0"))
        prop.requestedHandleTypes = cu.CUmemAllocationHandleType.CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR
        gran = ck(cu.cuMemGetAllocationGranularity(prop, cu.CUmemAllocationGranularity_flags.CU_MEM_ALLOC_GRANULARITY_MINIMUM))
        size = ((WIDTH * HEIGHT * 8 + gran - 1) // gran) * gran
        handle = ck(cu.cuMemCreate(size, prop, 0))
        fd = ck(cu.cuMemExportToShareableHandle(handle, cu.CUmemAllocationHandleType.CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0))
        r.bind_external_image(RT["DENOISED"], r.EXT_CUDA_VMM_FD, int(fd), size, 0, WIDTH * 8)
    except Exception as e:  # the driver refused: report why instead of failing the bench
        return {"unavailable": str(e)}
    rays = 0
    sync_all()
    t0 = time.perf_counter()
    for _ in range(steps):
        scene.update_nodes()
        nodes_ptr, nn = scene.nodes()
        C.memmove(pinned_nodes.data_ptr(), nodes_ptr, node_bytes)
        r.update_nodes_ptr(pinned_nodes.data_ptr(), nn)
        frame(tf)
        tf += 1
        c = r.counters()  # 24 bytes of ray counters come back every step
        rays += c["primary_rays"] + c["closest_hit_rays"] + c["any_hit_rays"]
    sync_all()
    secs = time.perf_counter() - t0
    r.unbind_external_image(RT["DENOISED"])
    os.close(int(fd))
    ck(cu.cuMemRelease(handle))
    return {"value": rays / secs / 1e6, "unit": "Mrays/s", "ms_per_step": 1e3 * secs / steps, "h2d_bytes_per_step": node_bytes + 784, "d2h_bytes_per_step": 24,
            "present": "RT_DENOISED written into imported external memory (gkcu_bind_external_image), no image copy"}


def workload_config(n, probes=False):
    name = "BASELINE configs[1]" if (WIDTH, HEIGHT) == (1920, 1080) else "SURVEY 8d config 5 (qx50.glb absent: declared stand-in scene)"
    return {"workload": f"{name}: RayTracingInOneWeekend procedural scene (RTIO.proc), {WIDTH}x{HEIGHT}, {SPP} spp per GPU "
                        f"({SPP * n} spp/frame), {BOUNCES} bounces, {'sky ' + os.path.basename(SKY_PATH) if SKY_PATH else 'procedural sky stand-in'}, steady state, temporal reproject + bilateral denoise",
            "scene": SCENE, "width": WIDTH, "height": HEIGHT, "spp_per_gpu": SPP, "spp_per_frame": SPP * n, "bounces": BOUNCES,
            "parallelism": f"spp-split x{n}, scene replica per GPU" if n > 1 else "single GPU",
            "l2": "working set (queue slots + hit records + per-pixel path state ~240 B/px, storage images 122 B/px: ~750 MB at 1080p) exceeds the 126 MB L2; "
                  "no explicit flush",
            "sky": f"Radiance .hdr equirect {os.path.basename(SKY_PATH)} (--sky, host.load_hdr)" if SKY_PATH else
                   "procedural equirect sky (host.make_sky 1024x512): the reference's assets/textures/river_road_2.hdr is reference content that neither "
                   "enters the repo nor travels to the GPU box; gkcu_set_sky takes any RGBA32F equirect and --sky PATH loads a Radiance .hdr",
            "probes": "voxelised + baked grid, bake dispatch every frame" if probes else "zero probe state"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--tracer", type=int, default=0, help="0 wavefront (default), 1 megakernel")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-strong", action="store_true", help="skip the config-5 strong-scaling block (4K, 64 spp split over the GPUs)")
    ap.add_argument("--cpu-leg", default=None, help=argparse.SUPPRESS)
    ap.add_argument("--sky", default=None, help="Radiance .hdr equirect used as sky 0 instead of the procedural stand-in")
    ap.add_argument("--probes", action="store_true",
                    help="voxelise the probe grid, bake one full sweep before timing and record the per-frame bake dispatch like "
                         "RayTraceBaseRenderer::PostRender; default: zero probe state (SURVEY 8d config 1/2)")
    ap.add_argument("--present", default="async", choices=["async", "blocking"],
                    help="e2e leg: async = double-buffered gkcu_readback_denoised_async overlapping the next frame; blocking = gkcu_readback")
    ap.add_argument("--reduce", default="peer-post", choices=["peer-post", "peer", "peer-hostbarrier", "nccl"],
                    help="multi-GPU: peer-post = every rank resolves, reprojects and composes its own strip over NVLink peer memory, history strips "
                         "mirrored to all ranks, nothing left to rank 0 alone (default); peer = fused reduce + resolve kernel over peer memory with "
                         "device-side barriers, post passes on rank 0; "
                         "peer-hostbarrier = same kernel bracketed by two 4-byte NCCL all-reduces; nccl = ncclReduce to rank 0, resolve there")
    ap.add_argument("--workload", default="config2", choices=["config2", "config5"],
                    help="config2: BASELINE configs[1], 1080p, 16 spp per GPU (weak scaling). config5: SURVEY 8d config 5, 3840x2160, 64 spp per frame "
                         "split over the GPUs (strong scaling)")
    args = ap.parse_args()
    global SKY_PATH
    SKY_PATH = args.sky
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup
    if args.cpu_leg:
        set_workload(args.workload, 1)
        return cpu_leg(args.cpu_leg)
    if args.impl == "reference":
        set_workload(args.workload, 1)
        return reference_arm(args)

    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: there is no CPU fallback (use --impl reference for the CPU port)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    out = run_ours(args, args.workload, args.steps, args.warmup, world, rank, local_rank, full=True)
    if args.workload == "config2" and not args.no_strong and 64 % world == 0:
        # The north-star multi-GPU configuration next to the weak-scaling headline: SURVEY 8d config 5, 3840x2160, 64 spp per frame SPLIT over
        # the GPUs (strong scaling: total work fixed), device-timed, a few frames. Reported as a block of the same line.
        st = run_ours(args, "config5", min(args.steps, 5), 3, world, rank, local_rank, full=False)
        if rank == 0:
            out["strong"] = {"workload": st["config"]["workload"], "scaling": "strong", "value": st["value"], "unit": "Mrays/s", "ms_per_step": st["ms_per_step"],
                             "steps": st["steps"], "warmup": st["warmup"], "spp_per_gpu": st["config"]["spp_per_gpu"], "width": 3840, "height": 2160,
                             "passes_ms_per_frame_rank0": st["passes_ms_per_frame_rank0"], "multi_gpu": st.get("multi_gpu")}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def set_workload(workload, world):
    global WIDTH, HEIGHT, SPP
    if workload == "config5":
        WIDTH, HEIGHT = 3840, 2160
        if 64 % world:
            raise SystemExit("config5 splits 64 spp: --gpus must divide 64")
        SPP = 64 // world
    else:
        WIDTH, HEIGHT, SPP = 1920, 1080, 16


def run_ours(args, workload, steps, warmup, world, rank, local_rank, full):
    """One measurement of `workload` on this rank's GPU: device-resident timing always; the end-to-end leg, the roofline objects and the CPU
    baseline only when `full`. Returns the JSON object on rank 0 (None elsewhere)."""
    import numpy as np  # noqa: F401
    import torch
    import torch.distributed as dist

    from gknextengine_b200 import RT, CudaPathTracingRenderer, GkcuFrameParams, host
    from gknextengine_b200._abi import GkNodeProxy

    set_workload(workload, world)
    args = argparse.Namespace(**vars(args))
    args.steps, args.warmup, args.workload = steps, warmup, workload
    scene = host.Scene(SCENE, samples=SPP * world, bounces=BOUNCES, taa=True, temporal_frames=16, denoiser=True)
    scene.update_nodes()  # steady state: combinedPrevTS = identity, MOTIONMOMENT stays 0
    sky = bench_sky(host)
    r = CudaPathTracingRenderer(local_rank)
    r.CreateSwapChain((WIDTH, HEIGHT))
    r.OnPostLoadScene(scene, sky)
    if args.probes:
        r.set_lights(scene)
        r.set_sky_sh(0, host.project_sh(sky))
        r.PreRender()
        r.probes_voxelize()
        ubo0 = scene.ubo(WIDTH, HEIGHT, 1)
        for f in range(120):
            r.PostRender(ubo0, f)
    ext = torch.cuda.ExternalStream(r.stream(), device=local_rank)
    params = GkcuFrameParams(temporal_frames=16, spp_world=world, spp_rank=rank, tracer=args.tracer, skip_post=1 if world > 1 else 0)

    radiance = None
    if world > 1:
        ptr, nbytes = r.radiance_buffer()

        class _Buf:
            __cuda_array_interface__ = {"shape": (nbytes // 4,), "typestr": "<f4", "data": (ptr, False), "version": 3}

        radiance = torch.as_tensor(_Buf(), device=f"cuda:{local_rank}")
        if args.reduce != "nccl":
            # ship every rank's buffers as cudaIpcMemHandles; each rank opens its peers' (NVLink peer access)
            mine = {what: r.peer_export(what)[0] for what in range(r.PEER_KINDS)}
            everyone = [None] * world
            dist.all_gather_object(everyone, mine)
            for pr in range(world):
                if pr == rank:
                    r.peer_import(pr, r.PEER_RADIANCE, ptr=r.peer_export(r.PEER_RADIANCE)[1])
                else:
                    r.peer_import(pr, r.PEER_RADIANCE, handle=everyone[pr][r.PEER_RADIANCE])
                for what in ((r.PEER_FLAGS,) if args.reduce == "peer" else (r.PEER_FLAGS,) + r.PEER_ACCUM if args.reduce == "peer-post" else ()):
                    if pr == rank:
                        r.peer_import(pr, what, ptr=r.peer_export(what)[1])
                    else:
                        r.peer_import(pr, what, handle=everyone[pr][what])
            for what in (r.PEER_DIRECT, r.PEER_SINGLE_DIFFUSE, r.PEER_SINGLE_SPECULAR, r.PEER_DENOISED):
                if rank == 0:
                    r.peer_import(0, what, ptr=r.peer_export(what)[1])
                else:
                    r.peer_import(0, what, handle=everyone[0][what])
        flag = torch.zeros(1, device=f"cuda:{local_rank}")
    reduce_ev = [torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)]
    reduce_ms = [0.0]

    def frame(total_frames):
        ubo = scene.ubo(WIDTH, HEIGHT, total_frames)
        r.PreRender()  # "TLAS Update" (RayTraceBaseRenderer.cpp:205-218): instance records + boxes from this frame's NodeProxy[], refit of the last full build (gkcu_set_tlas_refit)
        r.Render(ubo, params)
        if world > 1:
            reduce_ev[0].record(ext)
            if args.reduce == "nccl":
                with torch.cuda.stream(ext):
                    dist.reduce(radiance, dst=0, op=dist.ReduceOp.SUM)
                if rank == 0:
                    r.Resolve(ubo, params)
            elif args.reduce == "peer-post":
                r.ResolvePostStrip(ubo, params)  # barrier, strip resolve + strip reproject (history mirrored to all ranks) + strip compose into rank 0, barrier
            elif args.reduce == "peer":
                r.ResolveStrip(ubo, params)    # device-side barrier, peer loads of my strip from every rank, peer stores into rank 0's SINGLE_*, barrier
                if rank == 0:
                    r.run_post(ubo, params)
            else:
                with torch.cuda.stream(ext):
                    dist.all_reduce(flag)      # stream-ordered barrier: every rank's sums are complete
                r.ResolveStrip(ubo, params)
                with torch.cuda.stream(ext):
                    dist.all_reduce(flag)      # every strip has landed (and nobody still reads my radiance buffer)
                if rank == 0:
                    r.run_post(ubo, params)
            reduce_ev[1].record(ext)
        if args.probes:
            r.PostRender(ubo, total_frames)  # "hw-lightbake": 1/120 of the probe grid per frame (RayTraceBaseRenderer.cpp:225-274)
        return ubo

    def sync_all():
        r.BeforeNextFrame()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    # ---- device-resident timing (value). Its timed region carries the per-launch CUDA-event brackets the roofline object is computed from
    # (gkcu_set_kernel_timing(1); the library's default is off, and the end-to-end legs below run the default).
    r.set_kernel_timing(True)
    tf = 1
    for _ in range(args.warmup):
        frame(tf)
        tf += 1
    sync_all()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    rays = 0
    launches = 0
    kstats = {"trace_ms": 0.0, "shade_ms": 0.0, "tail_ms": 0.0, "primary_ms": 0.0, "trace_launches": 0, "trace_rays": 0, "tail_rays": 0}
    pass_ms = {"tlas_update_ms": 0.0, "rt_pass_ms": 0.0, "reproject_pass_ms": 0.0, "compose_pass_ms": 0.0}
    any_rays = closest_rays = primary_rays = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ext)
    for _ in range(args.steps):
        frame(tf)
        tf += 1
        c = r.counters()  # synchronises the stream: the frame is complete
        k = r.kernel_stats()
        t = r.timings()
        if world > 1:
            reduce_ms[0] += reduce_ev[0].elapsed_time(reduce_ev[1])
        rays += c["primary_rays"] + c["closest_hit_rays"] + c["any_hit_rays"]
        primary_rays += c["primary_rays"]
        closest_rays += c["closest_hit_rays"]
        any_rays += c["any_hit_rays"]
        launches += c["kernel_launches"]
        for key in kstats:
            kstats[key] += k[key]
        for key in pass_ms:
            pass_ms[key] += t[key]
    e1.record(ext)
    sync_all()
    clocks = sampler.stop() if rank == 0 else None
    ms = e0.elapsed_time(e1)
    # The timed region above carries CUDA-event brackets around every k_trace / k_shade launch (the per-kernel durations of the roofline
    # object); they cost ~4 us each. The same K frames without them (gkcu_set_kernel_timing(0)) are reported next to the headline.
    r.set_kernel_timing(False)
    frame(tf)
    tf += 1
    sync_all()
    n0, n1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0.record(ext)
    for _ in range(args.steps):
        frame(tf)
        tf += 1
        r.counters()
    n1.record(ext)
    sync_all()
    ms_plain = torch.tensor([n0.elapsed_time(n1)], dtype=torch.float64, device=f"cuda:{local_rank}")
    if world > 1:
        dist.all_reduce(ms_plain, op=dist.ReduceOp.MAX)
    ms_plain = float(ms_plain[0])
    stats = torch.tensor([ms, float(rays), float(launches)], dtype=torch.float64, device=f"cuda:{local_rank}")
    if world > 1:
        mx = stats.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms, rays, launches = float(mx[0]), float(sm[1]), float(sm[2])
    value = rays / (ms * 1e-3) / 1e6

    e2e_value, e2e_s, node_bytes, zero_copy = 0.0, 0.0, 0, None
    if full:
        # ---- end to end through the public API with host buffers (e2e)
        nodes_ptr, nn = scene.nodes()
        node_bytes = nn * C.sizeof(GkNodeProxy)
        pinned_nodes2 = [torch.empty(node_bytes, dtype=torch.uint8, pin_memory=True) for _ in range(2)]
        pinned_nodes = pinned_nodes2[0]
        pinned_out = [torch.empty(WIDTH * HEIGHT * 8, dtype=torch.uint8, pin_memory=True) for _ in range(2)]
        e2e_rays = 0

        def prepare(step):
            # the host side of a frame - Scene::UpdateNodes (Scene.cpp:410-492) and staging NodeProxy[] in pinned memory - runs while the
            # GPU renders the frame before it, the way the engine's CPU frame overlaps its GPU frame
            scene.update_nodes()
            ptr, _ = scene.nodes()
            C.memmove(pinned_nodes2[step & 1].data_ptr(), ptr, node_bytes)

        def e2e_steps(nsteps):
            nonlocal tf
            rays_total = 0
            prepare(0)
            for _step in range(nsteps):
                r.update_nodes_ptr(pinned_nodes2[_step & 1].data_ptr(), nn)  # H2D NodeProxy[]
                frame(tf)  # H2D UBO inside
                tf += 1
                if rank == 0:
                    # D2H of the frame's RT_DENOISED every step, double-buffered: the copy runs on the renderer's copy stream and overlaps
                    # the next frame's rt pass; the copy of the previous step is complete before its buffer is reused
                    out_buf = pinned_out[_step & 1]
                    if args.present == "async":
                        r.present_wait()
                        r.present_async(out_buf.data_ptr(), out_buf.numel())
                    else:
                        r.readback_into(RT["DENOISED"], out_buf.data_ptr(), out_buf.numel())
                if _step + 1 < nsteps:
                    prepare(_step + 1)
                c = r.counters()  # synchronises the stream: the frame is complete
                rays_total += c["primary_rays"] + c["closest_hit_rays"] + c["any_hit_rays"]
            if rank == 0:
                r.present_wait()
            return rays_total

        # W untimed steps of this very loop first: the first copies into freshly pinned host buffers are several times slower than the
        # steady state (7.5 ms against 0.4 ms for the 16.6 MB image on these boxes), and that belongs to the warm-up
        e2e_steps(args.warmup)
        sync_all()
        t0 = time.perf_counter()
        e2e_rays = e2e_steps(args.steps)
        sync_all()
        e2e_s = time.perf_counter() - t0
        est = torch.tensor([e2e_s, float(e2e_rays)], dtype=torch.float64, device=f"cuda:{local_rank}")
        if world > 1:
            mx = est.clone()
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            sm = est.clone()
            dist.all_reduce(sm, op=dist.ReduceOp.SUM)
            e2e_s, e2e_rays = float(mx[0]), float(sm[1])
        e2e_value = e2e_rays / e2e_s / 1e6
        if world == 1:
            zero_copy = zero_copy_leg(r, scene, frame, sync_all, args.steps, tf, pinned_nodes, node_bytes)
            tf += args.steps

    out = None
    if rank == 0:
        peak, peak_kind = load_peaks()
        out = {
            "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak" if args.workload == "config2" else "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(world, args.probes),
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "ms_per_step": 1e3 * e2e_s / args.steps, "h2d_bytes_per_step": node_bytes + 784,
                    "d2h_bytes_per_step": WIDTH * HEIGHT * 8,
                    "note": "library defaults (no per-launch event brackets, gkcu_set_kernel_timing = 0); value's timed region carries them for the roofline"} if full else None,
            "e2e_zero_copy": zero_copy,
            "gpu_launches": int(launches),
            "clocks": clocks,
            "tracer": "wavefront" if args.tracer == 0 else "megakernel",
            "rays_per_frame": {"primary": primary_rays / args.steps, "closest_hit": closest_rays / args.steps, "any_hit": any_rays / args.steps},
            "passes_ms_per_frame_rank0": {k: v / args.steps for k, v in pass_ms.items()},
            "tlas_update": dict(r.tlas_stats(), policy="refit of the last full build while the same instances are active, full SAH build otherwise (gkcu_set_tlas_refit); counts since the context was created"),
            "without_kernel_events": {"ms_per_step": ms_plain / args.steps, "value": value * ms / ms_plain, "unit": "Mrays/s",
                                      "note": "same K frames with gkcu_set_kernel_timing(0): no per-launch event brackets inside the rt pass"},
        }
        if world > 1:
            out["multi_gpu"] = {"exchange": {"peer-post": "strip-parallel tail over NVLink peer memory: every rank resolves (+7 halo rows), reprojects (+5) and composes its strip, mirrors its ACCUM_* strip into every rank, stores its RT_DENOISED strip into rank 0; device-side flag barriers, no collective call",
                                             "peer": "fused reduce+resolve kernel over NVLink peer memory (strips), device-side flag barriers, no collective call",
                                             "peer-hostbarrier": "fused reduce+resolve kernel over NVLink peer memory (strips) between two 4-byte NCCL all-reduces",
                                             "nccl": "ncclReduce(sum, fp32) to rank 0 + resolve on rank 0"}[args.reduce],
                                "exchange_resolve_post_ms_per_frame_rank0": reduce_ms[0] / args.steps,
                                "radiance_bytes_per_rank": 2 * WIDTH * HEIGHT * 16}
        # roofline of the dominant kernel: k_trace in the wavefront tracer, k_tail in the megakernel
        if args.tracer == 0 and kstats["trace_launches"] > 0:
            per_launch_ms = kstats["trace_ms"] / kstats["trace_launches"]
            per_launch_bytes = TRACE_BYTES_PER_RAY * kstats["trace_rays"] / kstats["trace_launches"]
            kernel, kms, share = "k_trace", kstats["trace_ms"], kstats["trace_ms"] / (ms if world == 1 else max(ms, 1e-9))
        else:
            per_launch_ms = kstats["tail_ms"] / args.steps
            per_launch_bytes = (PAIR_BYTES_CLOSEST * closest_rays + PAIR_BYTES_ANY * any_rays) / args.steps
            kernel, kms, share = "k_tail", kstats["tail_ms"], kstats["tail_ms"] / ms
        achieved = per_launch_bytes / (per_launch_ms * 1e-3) / 1e9 if per_launch_ms > 0 else 0.0
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                traffic = json.load(f).get(kernel)
        except Exception:
            pass
        out["roofline"] = {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak, "peak_kind": peak_kind, "unit": "GB/s",
                           "frac": achieved / peak, "traffic": traffic, "launches": kstats["trace_launches"] if kernel == "k_trace" else args.steps,
                           "avg_launch_ms": per_launch_ms, "share_of_step": share,
                           "kernels_ms_per_frame_rank0": {k: kstats[k] / args.steps for k in ("primary_ms", "trace_ms", "shade_ms", "tail_ms")}}
        rt_bytes = (PAIR_BYTES_CLOSEST * closest_rays + PAIR_BYTES_ANY * any_rays + PRIMARY_BYTES_PER_PIXEL * WIDTH * HEIGHT * args.steps * world) / world
        out["roofline_rt_pass"] = {"achieved": rt_bytes / (pass_ms["rt_pass_ms"] * 1e-3) / 1e9 if pass_ms["rt_pass_ms"] else 0.0, "unit": "GB/s",
                                   "note": "SURVEY 8d bytes (436 B/closest-hit ray, 100 B/any-hit ray, 38 B/pixel) over the whole rt pass, rank 0"}
        px = WIDTH * HEIGHT
        out["roofline_stencils"] = {
            "reproject": {"achieved": REPROJECT_BYTES_PER_PIXEL * px * args.steps / (pass_ms["reproject_pass_ms"] * 1e-3) / 1e9 if pass_ms["reproject_pass_ms"] else 0.0,
                          "bytes_per_pixel": REPROJECT_BYTES_PER_PIXEL},
            "denoise": {"achieved": DENOISE_BYTES_PER_PIXEL * px * args.steps / (pass_ms["compose_pass_ms"] * 1e-3) / 1e9 if pass_ms["compose_pass_ms"] else 0.0,
                        "bytes_per_pixel": DENOISE_BYTES_PER_PIXEL},
            "unit": "GB/s", "peak": peak}
        if full and world == 1 and not args.no_cpu_baseline:
            ubo = scene.ubo(WIDTH, HEIGHT, 1)
            p1 = GkcuFrameParams(temporal_frames=16, spp_world=1)
            bands, band_rows = 27, 40  # the whole frame: ~10 s of CPU work on a 16-core host
            crays, csecs, ckind, cwhat = cpu_sample(scene, sky, ubo, p1, bands, band_rows)
            out["cpu_baseline"] = {"value": crays / csecs / 1e6, "unit": "Mrays/s", "cores": os.cpu_count() or 1, "kind": ckind,
                                   "sample": f"{bands} bands x {band_rows} rows of the {WIDTH}x{HEIGHT} frame (rt pass), {csecs:.1f} s of CPU work; {cwhat}",
                                   "ms_per_frame_extrapolated": 1e3 * csecs * HEIGHT / (bands * band_rows)}
    r.close()
    return out if rank == 0 else None


if __name__ == "__main__":
    main()
