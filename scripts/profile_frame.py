"""Renders a few frames of the bench workload (for ncu): python scripts/profile_frame.py [frames] [tracer] [spp] [width] [height]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gknextengine_b200 import CudaPathTracingRenderer, GkcuFrameParams, host
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 2
tracer = int(sys.argv[2]) if len(sys.argv) > 2 else 0
spp = int(sys.argv[3]) if len(sys.argv) > 3 else 16
W = int(sys.argv[4]) if len(sys.argv) > 4 else 1920
H = int(sys.argv[5]) if len(sys.argv) > 5 else 1080
scene = host.Scene("RTIO.proc", samples=spp, bounces=8, taa=True, temporal_frames=16, denoiser=True)
scene.update_nodes()
r = CudaPathTracingRenderer(0)
r.set_kernel_timing(True)  # per-launch event brackets: kernel_stats() needs them
r.CreateSwapChain((W, H))
r.OnPostLoadScene(scene, host.make_sky(1024, 512))
p = GkcuFrameParams(temporal_frames=16, spp_world=1, tracer=tracer)
for f in range(1, frames + 1):
    r.PreRender()
    r.Render(scene.ubo(W, H, f), p)
    r.BeforeNextFrame()
print(r.timings(), r.kernel_stats(), r.counters())
r.close()
