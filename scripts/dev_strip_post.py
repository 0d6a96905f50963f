"""Bring-up of gkcu_render_resolve_post_strip with `world` contexts on one GPU: python scripts/dev_strip_post.py [world] [frames]"""
import os, sys, time
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gknextengine_b200 import RT, CudaPathTracingRenderer, GkcuFrameParams, host
world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 5
W, H = 256, 146
scene = host.Scene("RTIO.proc", samples=6, bounces=4, taa=True, temporal_frames=8, denoiser=True)
scene.update_nodes()
sky = host.make_sky(256, 128)
rs = []
for rank in range(world):
    r = CudaPathTracingRenderer(0)
    r.CreateSwapChain((W, H))
    r.OnPostLoadScene(scene, sky)
    rs.append(r)
exports = [{what: r.peer_export(what)[1] for what in range(r.PEER_KINDS)} for r in rs]
for r in rs:
    for rank in range(world):
        for what in (r.PEER_RADIANCE, r.PEER_FLAGS) + r.PEER_ACCUM:
            r.peer_import(rank, what, ptr=exports[rank][what])
    for what in (r.PEER_DIRECT, r.PEER_DENOISED):
        r.peer_import(0, what, ptr=exports[0][what])
def P(**kw):
    d = dict(temporal_frames=8, spp_world=world, skip_post=1)
    d.update(kw)
    return GkcuFrameParams(**d)
for f in range(1, frames + 1):
    ubo = scene.ubo(W, H, f, model_rotation_deg=(f / 60.0) * 40.0)
    t0 = time.perf_counter()
    for rank, r in enumerate(rs):
        r.Render(ubo, P(spp_rank=rank))
    t1 = time.perf_counter()
    for rank, r in enumerate(rs):
        r.ResolvePostStrip(ubo, P(spp_rank=rank))
        print(f"frame {f}: rank {rank} enqueued after {time.perf_counter() - t1:.3f} s", flush=True)
    for rank, r in enumerate(rs):
        try:
            r.BeforeNextFrame()
            print(f"frame {f}: rank {rank} synced, {time.perf_counter() - t0:.3f} s", flush=True)
        except Exception as e:
            print(f"frame {f}: rank {rank}: {e}", flush=True)
for r in rs:
    r.close()
