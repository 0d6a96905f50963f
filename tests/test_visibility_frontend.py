"""The visibility front end: "gpu cull" (Task.GpuCull.comp.slang:199-227) and the "visibility pass"
(Rast.VisibilityPass.{vert,frag}.slang, VulkanBaseRenderer.cpp:597-725).

CPU part: the oracle's rasteriser against an independent statement of visibility - the closest hit of a ray through the pixel centre on the
same fp16 vertices (they may only differ on silhouette pixels, where snapping to 1/256 pixel decides) - and the cull against hand-made cases.
GPU part: csrc/gk_raster.cu against the oracle, bit for bit: draw commands, GPUDrivenStat, RT_MINIGBUFFER, and a whole frame rendered from the
rasterised buffer (GkcuFrameParams.primary_visibility = 1) like FVisibilityBufferRayCasterV2 (Shading.slang:338-494)."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as O
from gknextengine_b200 import RT, GkcuFrameParams, host


def _scene(name, **kw):
    s = host.Scene(name, **kw)
    s.update_nodes()
    return s


def _cast_visibility(osc, ubo, W, H):
    out = np.zeros((H, W), np.uint32)
    osc.L.gko_cast_visibility.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
    osc.L.gko_cast_visibility(osc._h, C.byref(ubo), W, H, out.ctypes.data)
    return out


@pytest.mark.parametrize("name,size", [("CornellBox.proc", (320, 180)), ("RTIO.proc", (384, 216))])
def test_oracle_raster_agrees_with_centre_rays_away_from_silhouettes(name, size):
    W, H = size
    scene = _scene(name, samples=1, bounces=1, taa=False)
    osc = O.OracleScene(scene)
    ubo = scene.ubo(W, H, 1)
    far = np.ones((H, W), np.float32)  # cleared depth: nothing is occluded
    cmds, stats = osc.cull(ubo, far)
    assert stats["CulledCount"] == 0 and stats["ProcessedCount"] == int((cmds["instanceCount"] == 1).sum()) > 0
    vis = osc.visibility_pass(ubo, cmds, W, H)
    ray = _cast_visibility(osc, ubo, W, H)
    differ = vis != ray
    # a pixel where the two disagree must lie on an edge: some 4-neighbour holds a different triangle in the ray-cast image
    edge = np.zeros_like(differ)
    edge[1:, :] |= ray[1:, :] != ray[:-1, :]
    edge[:-1, :] |= ray[1:, :] != ray[:-1, :]
    edge[:, 1:] |= ray[:, 1:] != ray[:, :-1]
    edge[:, :-1] |= ray[:, 1:] != ray[:, :-1]
    assert differ.mean() < 0.02, f"{differ.mean():.4f} of the pixels differ between the rasteriser and the centre rays"
    assert not (differ & ~edge).any(), "the rasteriser and the centre rays disagree inside a triangle"
    # format: 15-bit (node + 1) over 17-bit triangle, 0 = nothing
    nodes, nn = scene.nodes()
    assert ((vis >> 17) <= nn).all() and (vis != 0).any()


def test_oracle_cull_frustum_occlusion_and_invisible_nodes():
    W, H = 160, 90
    scene = _scene("RTIO.proc", samples=1, bounces=1, taa=False)
    osc = O.OracleScene(scene)
    ubo = scene.ubo(W, H, 1)
    nodes, nn = scene.nodes()
    cmds, stats = osc.cull(ubo, np.ones((H, W), np.float32))
    drawn = cmds["instanceCount"] == 1
    assert 0 < drawn.sum() <= nn and stats["CulledCount"] == 0 and stats["VisibleCount"] == 0  # the shader never counts VisibleCount
    assert (cmds["firstInstance"][drawn] == np.nonzero(drawn)[0]).all()
    assert (cmds["indexCount"][~drawn] == 0).all() and (cmds["firstIndex"][~drawn] == 0).all()
    # a depth image of zeros is in front of everything: every batch whose 216 grid points all project into the image is culled
    cmds0, stats0 = osc.cull(ubo, np.zeros((H, W), np.float32))
    assert stats0["ProcessedCount"] == stats["ProcessedCount"] and 0 < stats0["CulledCount"] <= stats0["ProcessedCount"]
    assert stats0["CulledCount"] == drawn.sum() - (cmds0["instanceCount"] == 1).sum()
    assert stats0["CulledTriangleCount"] == int(cmds["indexCount"][drawn & (cmds0["instanceCount"] == 0)].sum()) // 3
    # an invisible node draws nothing and is not counted
    scene.set_node_visible(int(np.nonzero(drawn)[0][0]), False)
    scene.update_nodes()
    osc.update_nodes(scene)
    cmds1, stats1 = osc.cull(ubo, np.ones((H, W), np.float32))
    assert stats1["ProcessedCount"] == stats["ProcessedCount"] - 1 and cmds1["instanceCount"][np.nonzero(drawn)[0][0]] == 0


# ------------------------------------------------------------------------------------------------------------------- GPU
def _gpu_pair(scene, W, H):
    from gknextengine_b200 import CudaPathTracingRenderer
    r = CudaPathTracingRenderer(0)
    r.CreateSwapChain((W, H))
    r.OnPostLoadScene(scene)
    return r, O.OracleScene(scene)


@pytest.mark.gpu
@pytest.mark.parametrize("name,size", [("CornellBox.proc", (640, 360)), ("RTIO.proc", (1920, 1080))])
def test_cull_and_visibility_buffer_match_the_oracle(name, size):
    W, H = size
    scene = _scene(name, samples=1, bounces=1, taa=False)
    ubo = scene.ubo(W, H, 1)
    r, osc = _gpu_pair(scene, W, H)
    try:
        nodes, nn = scene.nodes()
        rng = np.random.default_rng(5)
        # last frame's depth: far everywhere except a few near rectangles, so some batches are occlusion-culled
        depth = np.ones((H, W), np.float32)
        for _ in range(12):
            x, y = int(rng.integers(0, W - W // 6)), int(rng.integers(0, H - H // 6))
            depth[y:y + H // 6, x:x + W // 6] = 0.0
        for d in (np.ones((H, W), np.float32), depth):
            r.upload_image(RT["PREV_DEPTHBUFFER"], d)
            r.cull(ubo)
            r.visibility_pass(ubo)
            cmds_cpu, stats_cpu = osc.cull(ubo, d)
            cmds_gpu = r.draw_commands(nn)
            assert (cmds_gpu == cmds_cpu).all()
            assert r.gpu_driven_stats() == stats_cpu
            vis_cpu = osc.visibility_pass(ubo, cmds_cpu, W, H)
            vis_gpu = r.readback(RT["MINIGBUFFER"])
            assert (vis_gpu == vis_cpu).all(), f"{(vis_gpu != vis_cpu).sum()} texels of RT_MINIGBUFFER differ"
            assert (vis_gpu != 0).mean() > 0.3
        assert stats_cpu["CulledCount"] > 0 or name == "CornellBox.proc"
    finally:
        r.close()


@pytest.mark.gpu
def test_frame_from_the_rasterised_visibility_buffer_matches_the_oracle():
    """Two frames with primary_visibility = 1: the second culls against the first frame's depth image."""
    from helpers import check_frame
    W, H = 640, 360
    scene = _scene("CornellBox.proc", samples=2, bounces=4, taa=True, denoiser=True)
    r, osc = _gpu_pair(scene, W, H)
    ofr = O.OracleFrame(W, H)
    try:
        p = GkcuFrameParams(temporal_frames=16, spp_world=1, primary_visibility=1)
        for frame in (1, 2):
            ubo = scene.ubo(W, H, frame)
            r.Render(ubo, p)
            ofr.render(osc, ubo, p)
            assert (r.readback(RT["MINIGBUFFER"]) == ofr.image(RT["MINIGBUFFER"])).all()
            check_frame(r, ofr)
            r.BeforeNextFrame()
            ofr.end_frame_copies()
        assert r.counters()["primary_rays"] < W * H // 4  # only depth-of-field re-casts remain; the centre rays are gone
    finally:
        r.close()


@pytest.mark.gpu
def test_visibility_pass_rejects_what_the_format_cannot_hold():
    from gknextengine_b200 import CudaPathTracingRenderer
    scene = _scene("CornellBox.proc", samples=1, bounces=1, taa=False)
    r = CudaPathTracingRenderer(0)
    try:
        r.CreateSwapChain((64, 64))
        r.OnPostLoadScene(scene)
        ubo = scene.ubo(64, 64, 1)
        with pytest.raises(RuntimeError, match="gpu cull"):
            r.visibility_pass(ubo)  # no draw commands yet
    finally:
        r.close()
