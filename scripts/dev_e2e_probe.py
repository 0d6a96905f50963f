"""Where the end-to-end loop of bench.py spends its host time: python scripts/dev_e2e_probe.py [steps]"""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from gknextengine_b200 import CudaPathTracingRenderer, GkcuFrameParams, host
from gknextengine_b200._abi import GkNodeProxy
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
W, H = 1920, 1080
scene = host.Scene("RTIO.proc", samples=16, bounces=8, taa=True, temporal_frames=16, denoiser=True)
scene.update_nodes()
r = CudaPathTracingRenderer(0)
r.CreateSwapChain((W, H))
r.OnPostLoadScene(scene, host.make_sky(1024, 512))
p = GkcuFrameParams(temporal_frames=16, spp_world=1, tracer=0)
_, nn = scene.nodes()
nb = nn * C.sizeof(GkNodeProxy)
pin = [torch.empty(nb, dtype=torch.uint8, pin_memory=True) for _ in range(2)]
out = [torch.empty(W * H * 8, dtype=torch.uint8, pin_memory=True) for _ in range(2)]
def prepare(k):
    scene.update_nodes(); ptr, _ = scene.nodes(); C.memmove(pin[k & 1].data_ptr(), ptr, nb)
def frame(tf):
    r.PreRender(); r.Render(scene.ubo(W, H, tf), p)
tf = 1
for _ in range(3):
    frame(tf); tf += 1
r.counters(); torch.cuda.synchronize()
for mode in ("none", "async", "async-h2d-first"):
    T = {k: 0.0 for k in ("upload", "frame", "present", "prepare", "counters")}
    prepare(0)
    if mode == "async-h2d-first": r.update_nodes_ptr(pin[0].data_ptr(), nn)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for s in range(steps):
        a = time.perf_counter()
        if mode != "async-h2d-first": r.update_nodes_ptr(pin[s & 1].data_ptr(), nn)
        b = time.perf_counter(); T["upload"] += b - a
        frame(tf); tf += 1
        a = time.perf_counter(); T["frame"] += a - b
        if mode == "async-h2d-first":
            prepare(s + 1); b = time.perf_counter(); T["prepare"] += b - a
            r.update_nodes_ptr(pin[(s + 1) & 1].data_ptr(), nn); a = time.perf_counter(); T["upload"] += a - b
        if mode != "none":
            r.present_wait(); r.present_async(out[s & 1].data_ptr(), out[s & 1].numel())
        b = time.perf_counter(); T["present"] += b - a
        if mode != "async-h2d-first":
            prepare(s + 1); a = time.perf_counter(); T["prepare"] += a - b; b = a
        r.counters(); a = time.perf_counter(); T["counters"] += a - b
    if mode != "none": r.present_wait()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(mode, "ms/step %.3f" % (1e3 * dt / steps), {k: round(1e3 * v / steps, 3) for k, v in T.items()})
# raw D2H of the image
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
src = torch.empty(W * H * 8, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize(); e0.record(); out[0].copy_(src, non_blocking=True); e1.record(); torch.cuda.synchronize()
print("raw D2H 16.6 MB: %.3f ms" % e0.elapsed_time(e1))
r.close()
