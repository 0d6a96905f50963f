import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from gknextengine_b200 import RT, CudaPathTracingRenderer, GkcuFrameParams, host
W, H = 1920, 1080
scene = host.Scene("RTIO.proc", samples=2, bounces=4, taa=True, denoiser=True); scene.update_nodes()
sky = host.make_sky(256, 128)
ubo = scene.ubo(W, H, 1)
outs = []
for k in range(3):
    r = CudaPathTracingRenderer(0); r.CreateSwapChain((W, H)); r.OnPostLoadScene(scene, sky)
    r.Render(ubo, GkcuFrameParams(temporal_frames=16, spp_world=1)); r.BeforeNextFrame()
    outs.append({n: r.readback(RT[n]) for n in ("SINGLE_DIFFUSE","SINGLE_SPECULAR","ALBEDO","NORMAL","OBJEDCTID_0","MOTIONVECTOR","MOTIONMOMENT","ACCUMLATE_DIFFUSE","ACCUMLATE_SPECULAR","ACCUMLATE_ALBEDO","DENOISED")})
    r.close()
for n in outs[0]:
    for k in (1, 2):
        a, b = outs[0][n], outs[k][n]
        ne = (a != b)
        if ne.ndim == 3: ne = ne.any(axis=2)
        if ne.any():
            ys, xs = np.nonzero(ne)
            print(n, k, "differs at", ne.sum(), "pixels; first", list(zip(ys[:5], xs[:5])), "y range", ys.min(), ys.max(), "x range", xs.min(), xs.max())
            y, x = ys[0], xs[0]
            print("   values", a[y, x], b[y, x])
print("done")
