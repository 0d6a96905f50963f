"""How many pixels of the bench frame have a dielectric primary hit (no early exit: the frame's longest paths): python scripts/dev_glass_pixels.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gknextengine_b200 import CudaPathTracingRenderer, GkcuFrameParams, RT, host
W, H = 1920, 1080
scene = host.Scene("RTIO.proc", samples=16, bounces=8, taa=True, temporal_frames=16, denoiser=True)
scene.update_nodes()
r = CudaPathTracingRenderer(0)
r.CreateSwapChain((W, H))
r.OnPostLoadScene(scene, host.make_sky(1024, 512))
r.PreRender()
r.Render(scene.ubo(W, H, 1), GkcuFrameParams(temporal_frames=16, spp_world=1))
r.BeforeNextFrame()
ids = r.readback(RT["OBJEDCTID_0"]).astype(np.int64).ravel()
nodes, nn = scene.nodes()
mats, nm = scene.materials()
model = np.array([mats[i].MaterialModel for i in range(nm)])
node_mat = np.array([nodes[i].matId[0] for i in range(nn)])
hit = (ids >= 0) & (ids < nn)
glass = np.zeros(ids.shape, bool)
glass[hit] = model[node_mat[ids[hit]]] == 2
print("pixels", ids.size, "primary hits", int(hit.sum()), "dielectric primary hits", int(glass.sum()), "= %.2f %%" % (100.0 * glass.sum() / ids.size))
print(r.counters(), r.kernel_stats())
r.close()
