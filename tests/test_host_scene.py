"""Host-side scene code (gknextengine_b200/host): fixture known-answers from SURVEY appendix C and the UBO construction."""
import ctypes as C

import numpy as np

from gknextengine_b200 import host
from gknextengine_b200._abi import GkNodeProxy


def nodes_array(scene):
    p, n = scene.nodes()
    return [p[i] for i in range(n)]


def mat(m):
    return np.array([[m.m[c][r] for c in range(4)] for r in range(4)], dtype=np.float64)  # math layout [row][col]


def test_rtiow_fixture_known_answers():
    s = host.Scene("RTIO.proc")
    c = s.counts()
    assert c == dict(nodes=1939, models=3, materials=1940, instanced_triangles=1984524)
    nodes = nodes_array(s)
    # first small sphere: libstdc++ mt19937(42) + uniform_real_distribution<float> (SceneList.cpp:140-141)
    t = mat(nodes[1].worldTS)[:3, 3]
    assert np.allclose(t, [-21.1443577, 0.949080229, -21.2831116], atol=1e-6)
    mats, nm = s.materials()
    models = [mats[i].MaterialModel for i in range(nm)]
    small = models[2:2 + 1935]
    assert (small.count(0), small.count(1), small.count(2)) == (1349, 392, 194)
    assert [n.modelId for n in nodes[:3]] == [0, 10, 10] and nodes[-1].modelId == 20
    assert all(n.instanceId == i for i, n in enumerate(nodes))


def test_cornell_fixture_known_answers():
    s = host.Scene("CornellBox.proc")
    assert s.counts() == dict(nodes=3, models=3, materials=7, instanced_triangles=1048)
    nodes = nodes_array(s)
    assert list(nodes[0].matId[:4]) == [1, 2, 3, 4] and nodes[1].matId[0] == 6 and nodes[2].matId[0] == 5
    mats, _ = s.materials()
    assert mats[4].MaterialModel == 4 and mats[4].Diffuse.x == 2000.0
    assert mats[6].MaterialModel == 5 and abs(mats[6].Fuzziness - 0.01) < 1e-9 and abs(mats[6].RefractionIndex - 1.45) < 1e-6
    d = s.desc()
    assert d.model_count == 3 and d.offset_count == 30 and d.vertex_count == 24 + 24 + 561
    # vertex material slots of the Cornell mesh (FProcModel.cpp:47-107): left wall slot 1, right wall slot 0
    assert d.vertices[0].tangentw & 0xFF == 1 and d.vertices[4].tangentw & 0xFF == 0 and d.vertices[20].tangentw & 0xFF == 3
    # tall box: scaled (1,2,1) and rotated about Y by 0.25 rad
    m = mat(nodes[2].worldTS)
    assert abs(m[1, 1] - 2.0) < 1e-6 and abs(m[0, 0] - np.cos(0.25)) < 1e-6 and abs(m[0, 2] - np.sin(0.25)) < 1e-6


def test_first_update_reports_motion_then_steady():
    s = host.Scene("CornellBox.proc")
    n0 = nodes_array(s)[0]
    c = mat(n0.combinedPrevTS)
    assert abs(c[1, 3] + 100.0) < 1e-4  # Node ctor: prevTransform = translate(0,-100,0) (Node.cpp:166-170)
    s.update_nodes()
    c = mat(nodes_array(s)[0].combinedPrevTS)
    assert np.array_equal(c, np.eye(4))


def test_ubo_construction():
    s = host.Scene("CornellBox.proc", samples=1, bounces=4, taa=False)
    u = s.ubo(640, 360, 1)
    P, MV = mat(u.Projection), mat(u.ModelView)
    f = 1.0 / np.tan(np.radians(40.0) / 2)
    assert abs(P[0, 0] - f / (640 / 360)) < 1e-5 and abs(P[1, 1] + f) < 1e-5  # y flipped for Vulkan
    assert abs(P[2, 2] - 2000.0 / (0.2 - 2000.0)) < 1e-6 and P[3, 2] == -1.0
    assert np.allclose(mat(u.ModelViewInverse) @ MV, np.eye(4), atol=1e-5)
    assert np.allclose(mat(u.ProjectionInverse) @ P, np.eye(4), atol=1e-4)
    assert np.allclose(mat(u.ViewProjection), P @ MV, atol=1e-4)
    assert np.allclose(mat(u.PrevViewProjection), mat(u.ViewProjection))  # first frame: prev == current
    assert (u.TotalFrames, u.NumberOfSamples, u.NumberOfBounces, u.MaxNumberOfBounces, u.HasSky, u.HasSun, u.TAA) == (1, 1, 4, 10, 0, 0, 0)
    assert C.sizeof(u) == 784
    # TAA: Halton(2,3) sub-pixel jitter in Projection[2][0..1] (Engine.cpp:771-778)
    s2 = host.Scene("CornellBox.proc", taa=True, temporal_frames=16)
    u2 = s2.ubo(640, 360, 3)
    hx = 0.125 - 0.5  # Halton base 2, index 4
    hy = (1 / 9 + 1 / 3) - 0.5  # Halton base 3, index 4 = 0.4444
    assert abs(u2.Projection.m[2][0] - hx / 640 * 2) < 1e-7 and abs(u2.Projection.m[2][1] - hy / 360 * 2) < 1e-7
    assert np.allclose(mat(u2.ViewProjectionUnJit), mat(u.Projection) @ mat(u2.ModelView), atol=1e-4)
    u3 = s2.ubo(640, 360, 4)
    assert np.allclose(mat(u3.PrevViewProjection), mat(u2.ViewProjection))


def test_mat4_inverse_restatement():
    rng = np.random.default_rng(0)
    for _ in range(20):
        m = rng.normal(size=(4, 4)).astype(np.float32)
        out = np.zeros((4, 4), dtype=np.float32)
        host.lib().gkhost_mat4_inverse(m.ctypes.data, out.ctypes.data)
        assert np.allclose(out.astype(np.float64) @ m.astype(np.float64), np.eye(4), atol=2e-3)


def test_lego_synth_workload_shape():
    s = host.Scene("LegoBricks.synth:3000")
    c = s.counts()
    assert c["nodes"] == 3001 and c["models"] == 17
    full = 12 + 16 * 6.25
    assert 0.8 * full * 3000 < c["instanced_triangles"] < 1.2 * full * 3000
    before = [mat(n.worldTS)[1, 3] for n in nodes_array(s)]
    s.jitter_instances(1, 0.01)
    s.update_nodes()
    after = nodes_array(s)
    moved = sum(1 for a, b in zip(before, after) if a != mat(b.worldTS)[1, 3])
    assert 0 < moved <= 30
    sv = s.perturb_model(5, 3)
    assert len(sv) % 4 == 0 and sv.dtype == np.uint16


def test_sky_is_deterministic_and_finite():
    a, b = host.make_sky(64, 32), host.make_sky(64, 32)
    assert np.array_equal(a, b) and np.isfinite(a).all() and a[..., :3].min() >= 0 and (a[..., 3] == 1).all()


def test_invisible_nodes_keep_their_proxy_with_the_visible_word_cleared():
    """Scene::UpdateNodesGpuDriven (Scene.cpp:443-492) records every drawable node, visible or not: instance i of the TLAS is Nodes[i], so hiding
    a node must not shift the others - its proxy stays, with visible = 0 (the TLAS mask hides it, TopLevelAccelerationStructure.cpp:83-111)."""
    scene = host.Scene("CornellBox.proc")
    scene.update_nodes()
    p, n = scene.nodes()
    before = [(p[i].instanceId, p[i].modelId, p[i].visible) for i in range(n)]
    scene.set_node_visible(1, False)
    scene.update_nodes()
    p, n2 = scene.nodes()
    after = [(p[i].instanceId, p[i].modelId, p[i].visible) for i in range(n2)]
    assert n2 == n and after[1] == (before[1][0], before[1][1], 0) and after[0] == before[0] and after[2] == before[2]
    assert scene.changed()[1] == 1  # exactly one proxy differs from the previous frame
    scene.set_node_visible(1, True)
    scene.update_nodes()
    p, _ = scene.nodes()
    assert p[1].visible == 1


def test_provoking_vertex_index_buffers():
    """Scene::RebuildMeshBuffer passes every section through meshopt_generateProvokingIndexBuffer (Scene.cpp:224-250; meshoptimizer is not in
    the mount, the published algorithm is restated in host/gk_scene.cpp). Properties of the three buffers it yields: every triangle of
    GPUScene.Indices is a ROTATION of the authored triangle; Reorders[provoke[i]] is that same vertex; the provoking (first) index of every
    triangle is unique in its section, which is what lets a flat-interpolated vertex attribute carry a per-triangle value."""
    for name in ("CornellBox.proc", "RTIO.proc", "LegoBricks.synth:50"):
        s = host.Scene(name)
        d = s.desc()
        prim = np.ctypeslib.as_array(d.indices, shape=(d.index_count,))
        provoke, reorders = s.raster_buffers()
        assert len(provoke) == len(prim)
        for m in range(d.model_count):
            off = [d.offsets[m * 10 + k] for k in range(10)]
            authored = s.model_indices(m)
            done = 0
            for o in off:
                if o.indexCount == 0:
                    continue
                tri = prim[o.indexOffset:o.indexOffset + o.indexCount].reshape(-1, 3)
                src = authored[done:done + o.indexCount].reshape(-1, 3)
                done += o.indexCount
                rot = [(tri == np.roll(src, -k, axis=1)).all(axis=1) for k in range(3)]
                assert (rot[0] | rot[1] | rot[2]).all(), "a triangle is not a rotation of the authored one"
                pv = provoke[o.indexOffset:o.indexOffset + o.indexCount]
                assert (reorders[o.reorderOffset + pv] - o.vertexOffset == prim[o.indexOffset:o.indexOffset + o.indexCount]).all()
                first = pv.reshape(-1, 3)[:, 0]
                assert len(np.unique(first)) == len(first), "provoking vertices must be unique per section"
                assert pv.max() < len(reorders) - o.reorderOffset and len(reorders) - o.reorderOffset <= o.vertexCount + o.indexCount // 3 + (len(reorders) - o.reorderOffset)
            assert done == len(authored)
    # known answer, worked by hand from the algorithm: two triangles sharing an edge, 4 vertices, valence = [1, 2, 2, 1].
    #   tri (0,1,2): all corners unseen, a = 0 has the least valence -> stays, slot 0 = vertex 0; valence -> [0, 1, 1, 1]
    #   tri (2,1,3): all corners unseen with valence 1 each; the tie goes to b (the "bca" test comes first) -> (1,3,2), slot 1 = vertex 1
    #   second pass over the non-provoking corners: vertex 1 already has slot 1; vertex 2 -> slot 2; vertex 3 -> slot 3
    #   => provoke = [0,1,2, 1,3,2], reorder = [0,1,2,3]
    import ctypes as C
    L = host.lib()
    L.gkhost_provoking_index_buffer.restype = C.c_uint32
    L.gkhost_provoking_index_buffer.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32]
    idx = np.array([0, 1, 2, 2, 1, 3], np.uint32)
    dst, reo = np.zeros(6, np.uint32), np.zeros(4 + 2, np.uint32)
    n = L.gkhost_provoking_index_buffer(dst.ctypes.data, reo.ctypes.data, idx.ctypes.data, 6, 4)
    assert n == 4 and dst.tolist() == [0, 1, 2, 1, 3, 2] and reo[:4].tolist() == [0, 1, 2, 3]
    # a fan around vertex 0: the shared vertex can provoke only once, the later triangles rotate to a fresh corner
    idx = np.array([0, 1, 2, 0, 2, 3, 0, 3, 4], np.uint32)
    dst, reo = np.zeros(9, np.uint32), np.zeros(5 + 3, np.uint32)
    n = L.gkhost_provoking_index_buffer(dst.ctypes.data, reo.ctypes.data, idx.ctypes.data, 9, 5)
    tri = reo[dst].reshape(3, 3)
    assert n == 5 and len(set(dst.reshape(3, 3)[:, 0].tolist())) == 3
    assert tri.tolist() == [[1, 2, 0], [2, 3, 0], [3, 4, 0]]  # each time b is the unseen corner of least remaining valence (ties go to b)
