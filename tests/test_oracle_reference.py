"""Pins the oracle's intersector against the reference's own vendored tinybvh (compiled from /root/reference into
oracle/_ref/liboracle_tinybvh.so, same oracle source with -DGKO_TINYBVH) and checks frame-level invariants of the oracle."""
import os

import numpy as np
import pytest

import oracle_lib as O
from gknextengine_b200 import RT, GkcuFrameParams, host
from helpers import random_rays

need_ref = pytest.mark.skipif(not os.path.exists(O.TINYBVH), reason="oracle/_ref not built (reference not mounted)")


@need_ref
@pytest.mark.parametrize("name,center", [("CornellBox.proc", (0, 2.7, 0)), ("RTIO.proc", (0, 3, 0))])
def test_native_intersector_matches_tinybvh(name, center):
    s = host.Scene(name)
    s.update_nodes()
    a, b = O.OracleScene(s), O.OracleScene(s, lib_path=O.TINYBVH)
    assert O.load(O.TINYBVH).gko_intersector_name() == b"tinybvh-1.3.8"
    rays = random_rays(20000, seed=3, center=center)
    ha, hb = a.trace(rays), b.trace(rays)
    # coplanar surfaces of different instances (the Cornell tall box stands on the floor) are genuine ties: the oracle commits
    # the lower instance id, tinybvh whichever it meets first. Anything that is not a tie must agree.
    differ = ha["inst"] != hb["inst"]
    tie = differ & (ha["inst"] != 0xFFFFFFFF) & (hb["inst"] != 0xFFFFFFFF) & (np.abs(ha["t"] - hb["t"]) <= 1e-4 * np.maximum(1.0, ha["t"]))
    assert (differ & ~tie).mean() <= 1e-4
    same = ha["inst"] == hb["inst"]
    hit = same & (ha["inst"] != 0xFFFFFFFF)
    assert (ha["prim"][hit] != hb["prim"][hit]).mean() <= 2e-3  # shared edges: Moeller-Trumbore vs watertight pick either side
    assert np.abs(ha["t"][hit] - hb["t"][hit]).max() < 2e-3
    oa, ob = a.trace(rays, any_hit=True), b.trace(rays, any_hit=True)
    assert (oa["inst"] != ob["inst"]).mean() <= 1e-4


@need_ref
def test_primary_ids_match_tinybvh_on_config1():
    """BASELINE config 1 (Cornell box 640x360): primary-hit instance ids, <= 1e-4 mismatching (triangle edges)."""
    s = host.Scene("CornellBox.proc", samples=1, bounces=4, taa=False)
    s.update_nodes()
    p = GkcuFrameParams(temporal_frames=16, spp_world=1)
    ubo = s.ubo(640, 360, 1)
    imgs = []
    for path in (O.NATIVE, O.TINYBVH):
        sc, fr = O.OracleScene(s, lib_path=path), O.OracleFrame(640, 360, lib_path=path)
        fr.render(sc, ubo, p, run_post=False)
        imgs.append((fr.image(RT["OBJEDCTID_0"]).copy(), fr.image(RT["EXT_PRIMITIVEID"]).copy()))
    assert (imgs[0][0] != imgs[1][0]).mean() <= 1e-4
    assert (imgs[0][1] != imgs[1][1]).mean() <= 1e-3


def test_oracle_frame_invariants():
    s = host.Scene("CornellBox.proc", samples=2, bounces=4, taa=False)
    s.update_nodes()
    W, H = 160, 90
    sc, fr = O.OracleScene(s), O.OracleFrame(W, H)
    p = GkcuFrameParams(temporal_frames=16, spp_world=1)
    ubo = s.ubo(W, H, 1)
    fr.render(sc, ubo, p)
    ids = fr.image(RT["OBJEDCTID_0"]).copy()
    sd = fr.image(RT["SINGLE_DIFFUSE"]).astype(np.float32)
    c = fr.counters()
    assert c["primary"] == W * H and c["any"] == 0 and c["closest"] >= 2 * (ids != 65535).sum() * 0.9
    assert set(np.unique(ids)) <= {0, 1, 2, 65535}
    assert (sd[ids == 65535] == 0).all()  # no sky in the Cornell box
    prim = fr.image(RT["EXT_PRIMITIVEID"])
    light = (ids == 0) & (prim >= 10) & (prim < 12)  # the two light triangles
    assert light.sum() > 50 and (sd[light][:, :3] == 2000.0).all()
    # determinism and frame seeding
    fr2 = O.OracleFrame(W, H)
    fr2.render(sc, ubo, p)
    assert np.array_equal(fr2.image(RT["SINGLE_DIFFUSE"]), fr.image(RT["SINGLE_DIFFUSE"]))
    fr2.render(sc, s.ubo(W, H, 2), p)
    assert not np.array_equal(fr2.image(RT["SINGLE_DIFFUSE"]), fr.image(RT["SINGLE_DIFFUSE"]))
    # after the copy pass PREV == ACCUM and OBJECTID_1 == OBJECTID_0
    assert np.array_equal(fr.image(RT["SINGLE_PREV_DIFFUSE"]), fr.image(RT["ACCUMLATE_DIFFUSE"]))
    assert np.array_equal(fr.image(RT["OBJEDCTID_1"]), fr.image(RT["OBJEDCTID_0"]))


def test_reproject_properties():
    W, H = 48, 32
    fr = O.OracleFrame(W, H)
    rng = np.random.default_rng(5)
    cur = rng.uniform(0, 4, (H, W, 4)).astype(np.float16)
    prev = rng.uniform(0, 4, (H, W, 4)).astype(np.float16)
    fr.image(RT["SINGLE_DIFFUSE"])[:] = cur
    fr.image(RT["SINGLE_PREV_DIFFUSE"])[:] = prev
    # progressive: exact lerp(prev, cur, 1/512) on all four channels (Process.ReProject.comp.slang:95-101)
    fr.reproject(GkcuFrameParams(progressive=1, temporal_frames=16))
    expect = (prev.astype(np.float32) + (cur.astype(np.float32) - prev.astype(np.float32)) * np.float32(1 / 512)).astype(np.float16)
    assert np.array_equal(fr.image(RT["ACCUMLATE_DIFFUSE"]), expect)
    # a miss (object id 65535) never uses history: ACCUM = current
    fr.image(RT["OBJEDCTID_0"])[:] = 65535
    fr.reproject(GkcuFrameParams(progressive=0, temporal_frames=16))
    assert np.array_equal(fr.image(RT["ACCUMLATE_DIFFUSE"]), cur)
    # static uniform scene: history valid, constant colour stays constant
    fr.image(RT["OBJEDCTID_0"])[:] = 3
    fr.image(RT["OBJEDCTID_1"])[:] = 3
    fr.image(RT["NORMAL"])[:] = np.float16([0, 1, 0, 0.5])
    fr.image(RT["SINGLE_DIFFUSE"])[:] = np.float16([0.5, 0.25, 0.125, 1])
    fr.image(RT["SINGLE_PREV_DIFFUSE"])[:] = np.float16([0.5, 0.25, 0.125, 1])
    fr.reproject(GkcuFrameParams(progressive=0, temporal_frames=16))
    inner = fr.image(RT["ACCUMLATE_DIFFUSE"])[4:-4, 4:-4, :3].astype(np.float32)
    assert np.abs(inner - np.float32([0.5, 0.25, 0.125])).max() < 1e-3
    # MOTIONMOMENT > 0 disables history
    fr.image(RT["MOTIONMOMENT"])[:] = 2
    fr.image(RT["SINGLE_PREV_DIFFUSE"])[:] = 0
    fr.reproject(GkcuFrameParams(progressive=0, temporal_frames=16))
    assert np.array_equal(fr.image(RT["ACCUMLATE_DIFFUSE"]), fr.image(RT["SINGLE_DIFFUSE"]))


def test_denoise_properties():
    W, H = 40, 24
    s = host.Scene("CornellBox.proc", denoiser=True)
    ubo = s.ubo(W, H, 1)
    fr = O.OracleFrame(W, H)
    fr.image(RT["ACCUMLATE_DIFFUSE"])[:] = np.float16([0.5, 0.5, 0.5, 1])
    fr.image(RT["ACCUMLATE_ALBEDO"])[:] = np.float16([0.5, 1.0, 0.25, 1])
    fr.image(RT["ACCUMLATE_SPECULAR"])[:] = np.float16([0.125, 0, 0, 1])
    fr.denoise(ubo)
    out = fr.image(RT["DENOISED"]).astype(np.float32)
    assert (out[..., 3] == 1).all()
    inner = out[6:-6, 6:-6, :3]
    assert np.ptp(inner.reshape(-1, 3), axis=0).max() == 0  # uniform input -> uniform output away from the border
    # filter off: plain modulate + tonemap, identical everywhere including the border
    s2 = host.Scene("CornellBox.proc", denoiser=False)
    fr.denoise(s2.ubo(W, H, 1))
    out2 = fr.image(RT["DENOISED"]).astype(np.float32)
    assert np.ptp(out2.reshape(-1, 4), axis=0).max() == 0
    # the 0.001 fetch bias of the filtered path makes it slightly brighter than the unfiltered one
    assert (inner[0, 0] >= out2[0, 0, :3]).all()


REF_SH = os.path.join(os.path.dirname(O.TINYBVH), "libref_sh.so")


@pytest.mark.skipif(not os.path.exists(REF_SH), reason="oracle/_ref not built (reference not mounted)")
def test_sample_ibl_rough_is_pinned_to_the_references_own_evaluate_sh():
    """The oracle's SampleIBLRough (restating Shading.slang:146-192) against the reference's EvaluateSH / SampleIBL compiled from
    assets/shaders/glsl_legacy/common/SampleIBL.glsl where it lies (oracle/ref_sh.cpp): random coefficients, directions and sky
    rotations: bit-identical (sin / cos of the rotation angle are the pinned kernels of include/gk_detmath.h on both sides)."""
    import ctypes as C
    from gknextengine_b200 import host
    from gknextengine_b200._abi import GkSphericalHarmonics
    ref = C.CDLL(REF_SH)
    ref.ref_sample_ibl_sh.argtypes = [C.POINTER(C.c_float * 27), C.POINTER(C.c_float * 3), C.c_float, C.POINTER(C.c_float * 4)]
    scene = host.Scene("CornellBox.proc", samples=1, bounces=4)
    scene.update_nodes()
    osc = O.OracleScene(scene)
    rng = np.random.default_rng(5)
    worst = 0.0
    for trial in range(200):
        co = rng.normal(size=(3, 9)).astype(np.float32) * np.float32(3.0)
        sh = GkSphericalHarmonics()
        for c in range(3):
            for i in range(9):
                sh.coefficients[c][i] = float(co[c, i])
        osc.set_sky_sh(0, sh)
        d = rng.normal(size=3).astype(np.float32)
        d /= np.linalg.norm(d)
        rot = float(rng.uniform(0, 2)) if trial % 4 else 0.0
        a, b = (C.c_float * 4)(), (C.c_float * 4)()
        osc.L.gko_sample_ibl_rough(osc._h, 0, (C.c_float * 3)(*d), rot, a)
        ref.ref_sample_ibl_sh((C.c_float * 27)(*co.reshape(-1)), (C.c_float * 3)(*d), rot, b)
        a, b = np.array(a), np.array(b)
        scale = np.abs(co).sum(axis=1).max()
        worst = max(worst, float(np.abs(a[:3] - b[:3]).max() / scale))
        assert a[3] == b[3] == 1.0
    assert worst == 0.0


REF_CF = os.path.join(os.path.dirname(O.TINYBVH), "libref_const_func.so")


@pytest.mark.skipif(not os.path.exists(REF_CF), reason="oracle/_ref not built (reference not mounted)")
def test_sampling_rng_and_tonemap_library_is_pinned_to_the_references_source():
    """SURVEY 8a rows a19 / a25. The reference's own Const_Func.slang functions - extracted by line range where the file lies and compiled as
    C++ (oracle/ref_const_func.cpp, oracle/Makefile) - against the oracle's restatements on random inputs:
    pcg4d, uint_to_float, RandomFloat, RandomFloat2, Schlick, AlignWithNormal (ONB + to_world), GT tonemapper, ST2084, RandomInUnitDisk,
    RandomInCone, RandomInHemiSphere1, ggxSampling: all bit-exact, and the RNG state after each call identical. The intrinsics the text calls
    (sin, cos, pow) are the pinned kernels of include/gk_detmath.h on both sides - a GPU leaves their rounding to the driver."""
    import ctypes as C
    ref, L = C.CDLL(REF_CF), O.load()
    f3, f2, u4 = C.c_float * 3, C.c_float * 2, C.c_uint32 * 4
    for lib, pre in ((ref, "ref_"), (L, "gko_")):
        getattr(lib, pre + "schlick").restype = C.c_float
        getattr(lib, pre + "schlick").argtypes = [C.c_float, C.c_float]
        getattr(lib, pre + "uint_to_float").restype = C.c_float
        getattr(lib, pre + "uint_to_float").argtypes = [C.c_uint32]
        getattr(lib, pre + "random_float").restype = C.c_float
        getattr(lib, pre + "random_float").argtypes = [C.POINTER(u4)]
        getattr(lib, pre + "random_float2").argtypes = [C.POINTER(u4), C.POINTER(f2)]
        getattr(lib, pre + "random_in_unit_disk").argtypes = [C.POINTER(u4), C.POINTER(f2)]
        getattr(lib, pre + "random_in_cone").argtypes = [C.POINTER(u4), C.c_float, C.POINTER(f3)]
        getattr(lib, pre + "random_in_hemisphere1").argtypes = [C.POINTER(u4), C.POINTER(f3)]
        getattr(lib, pre + "ggx_sampling").argtypes = [C.POINTER(u4), C.c_float, C.POINTER(f3), C.POINTER(f3)]
        getattr(lib, pre + "align_with_normal").argtypes = [C.POINTER(f3), C.POINTER(f3), C.POINTER(f3)]
        getattr(lib, pre + "st2084").argtypes = [C.POINTER(f3), C.POINTER(f3)]
        getattr(lib, pre + "gt_tonemap").argtypes = [C.POINTER(f3), C.POINTER(f3)]
        getattr(lib, pre + "pcg4d").argtypes = [C.POINTER(u4)]
    bits = lambda a: np.array(a, dtype=np.float32).view(np.uint32).tolist()
    rng = np.random.default_rng(2024)
    worst = 0.0
    for trial in range(3000):
        seed = [int(x) for x in rng.integers(0, 2 ** 32, 4)]
        if trial < 4:
            seed = [[3, 5, 7, 0], [0, 0, 0, 0], [0xFFFFFFFF] * 4, [1919, 1079, 256, 0]][trial]
        # integer / exact functions
        a, b = u4(*seed), u4(*seed)
        ref.ref_pcg4d(a); L.gko_pcg4d(b)
        assert list(a) == list(b)
        assert bits([ref.ref_uint_to_float(seed[0])]) == bits([L.gko_uint_to_float(seed[0])])
        a, b = u4(*seed), u4(*seed)
        assert bits([ref.ref_random_float(a)]) == bits([L.gko_random_float(b)]) and list(a) == list(b)
        a, b, oa, ob = u4(*seed), u4(*seed), f2(), f2()
        ref.ref_random_float2(a, oa); L.gko_random_float2(b, ob)
        assert bits(oa) == bits(ob) and list(a) == list(b)
        c, ior = float(rng.uniform(0, 1)), float(rng.uniform(1.0, 2.5))
        assert bits([ref.ref_schlick(c, ior)]) == bits([L.gko_schlick(c, ior)])
        n = rng.normal(size=3); n /= np.linalg.norm(n)
        if trial % 50 == 0:
            n = np.array([0.0, 0.0, -1.0 if trial % 100 else 1.0])  # the ONB singularity side
        ray = rng.normal(size=3)
        oa, ob = f3(), f3()
        ref.ref_align_with_normal(f3(*ray), f3(*n), oa); L.gko_align_with_normal(f3(*ray), f3(*n), ob)
        assert bits(oa) == bits(ob)
        col = np.abs(rng.normal(size=3)) * (10.0 ** rng.uniform(-3, 1.5))
        ref.ref_gt_tonemap(f3(*col), oa); L.gko_gt_tonemap(f3(*col), ob)
        assert bits(oa) == bits(ob)
        ref.ref_st2084(f3(*(col * 100)), oa); L.gko_st2084(f3(*(col * 100)), ob)
        assert bits(oa) == bits(ob)
        # samplers with a trigonometric step
        for name, extra, out_t in (("random_in_unit_disk", (), f2), ("random_in_hemisphere1", (), f3),
                                   ("random_in_cone", (C.c_float(0.99999048),), f3), ("random_in_cone", (C.c_float(float(rng.uniform(-1, 1))),), f3)):
            a, b, oa, ob = u4(*seed), u4(*seed), out_t(), out_t()
            getattr(ref, "ref_" + name)(a, *extra, oa); getattr(L, "gko_" + name)(b, *extra, ob)
            assert list(a) == list(b)
            worst = max(worst, float(np.abs(np.array(oa) - np.array(ob)).max()))
        rough = float(rng.uniform(0.0, 1.0))
        a, b, oa, ob = u4(*seed), u4(*seed), f3(), f3()
        ref.ref_ggx_sampling(a, rough, f3(*n), oa); L.gko_ggx_sampling(b, rough, f3(*n), ob)
        assert list(a) == list(b)
        worst = max(worst, float(np.abs(np.array(oa) - np.array(ob)).max()))
    assert worst == 0.0, worst


@pytest.mark.skipif(not os.path.exists(REF_CF), reason="oracle/_ref not built (reference not mounted)")
def test_probe_colour_format_and_samplers_are_pinned_to_the_references_source():
    """AmbientCube.slang:58-223 (packRGB10A2, unpackRGB10A2, LerpPackedColorAlt, sampleAmbientCubeHL2_DI and
    _Full) compiled from the reference's file as C++ against the oracle's restatements: bit-exact on every 10-bit code, on random
    colours including out-of-range and negative ones, on random temporal lerps, and on random cubes and normals."""
    import ctypes as C
    ref, L = C.CDLL(REF_CF), O.load()
    f4, f3, c14 = C.c_float * 4, C.c_float * 3, C.c_uint32 * 14
    for lib, pre in ((ref, "ref_"), (L, "gko_")):
        getattr(lib, pre + "pack_rgb10a2").restype = C.c_uint32
        getattr(lib, pre + "pack_rgb10a2").argtypes = [C.POINTER(f4)]
        getattr(lib, pre + "unpack_rgb10a2").argtypes = [C.c_uint32, C.POINTER(f4)]
        getattr(lib, pre + "lerp_packed_color_alt").restype = C.c_uint32
        getattr(lib, pre + "lerp_packed_color_alt").argtypes = [C.c_uint32, C.POINTER(f4), C.c_float]
        getattr(lib, pre + "sample_ambient_cube").argtypes = [C.c_int, C.POINTER(c14), C.POINTER(f3), C.POINTER(f4)]
    bits = lambda a: np.array(a, dtype=np.float32).view(np.uint32).tolist()
    for code in range(1024):
        p = code | ((1023 - code) << 10) | (((code * 7) & 0x3FF) << 20) | ((code & 3) << 30)
        a, b = f4(), f4()
        ref.ref_unpack_rgb10a2(p, a); L.gko_unpack_rgb10a2(p, b)
        assert bits(a) == bits(b)
    rng = np.random.default_rng(77)
    for trial in range(4000):
        col = rng.normal(size=4) * (10.0 ** rng.uniform(-2, 3.2))
        assert ref.ref_pack_rgb10a2(f4(*col)) == L.gko_pack_rgb10a2(f4(*col))
        c0 = int(rng.integers(0, 2 ** 32))
        t = 0.125 if trial % 2 else float(rng.uniform(0, 1))
        assert ref.ref_lerp_packed_color_alt(c0, f4(*np.abs(col)), t) == L.gko_lerp_packed_color_alt(c0, f4(*np.abs(col)), t)
        cube = c14(*[int(x) for x in rng.integers(0, 2 ** 32, 14)])
        n = rng.normal(size=3)
        if trial % 3 == 0:
            n /= np.linalg.norm(n)
        if trial % 97 == 0:
            n[:] = 0  # sum of weights 0: the samplers' (sum > 0) guard
        for which in (0, 2):
            a, b = f4(), f4()
            ref.ref_sample_ambient_cube(which, cube, f3(*n), a); L.gko_sample_ambient_cube(which, cube, f3(*n), b)
            assert bits(a) == bits(b), (which, list(a), list(b))


REF_GF = os.path.join(os.path.dirname(O.TINYBVH), "libref_general_func.so")


@pytest.mark.skipif(not os.path.exists(REF_GF), reason="oracle/_ref not built (reference not mounted)")
def test_hit_attribute_reconstruction_is_pinned_to_the_references_source():
    """Common.get_material_dataV2 (with UnpackVertexV2 inside), FetchMaterialId, mikkTSpace and
    calculateWeight compiled from the reference's files as C++ (oracle/ref_general_func.cpp) against the oracle, on the real vertex / index /
    node buffers of both procedural scenes: every output float bit-exact (world-space tangents to 1 ulp-scale 2e-7: the oracle does not
    use them, it normalises them separately)."""
    import ctypes as C
    from gknextengine_b200 import host
    ref, L = C.CDLL(REF_GF), O.load()
    f18, f3, f4 = C.c_float * 18, C.c_float * 3, C.c_float * 4
    ref.ref_get_material_data.restype = C.c_uint32
    ref.ref_get_material_data.argtypes = [C.c_uint32, C.c_uint32, C.POINTER(f3), C.POINTER(f3), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(f18)]
    L.gko_get_material_data.restype = C.c_uint32
    L.gko_get_material_data.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(f3), C.POINTER(f3), C.POINTER(f18)]
    ref.ref_fetch_material_id.restype = L.gko_fetch_material_id.restype = C.c_uint32
    ref.ref_fetch_material_id.argtypes = [C.c_void_p, C.c_uint32]
    L.gko_fetch_material_id.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
    ref.ref_mikktspace.argtypes = L.gko_mikktspace.argtypes = [C.POINTER(f3), C.POINTER(f3), C.POINTER(f4), C.POINTER(f3)]
    ref.ref_calculate_weight.restype = L.gko_calculate_weight.restype = C.c_float
    ref.ref_calculate_weight.argtypes = L.gko_calculate_weight.argtypes = [C.c_float, C.c_int, C.c_int, C.POINTER(f3), C.POINTER(f3)]
    bits = lambda a: np.array(a, dtype=np.float32).view(np.uint32)
    rng = np.random.default_rng(31)
    for name in ("CornellBox.proc", "RTIO.proc"):
        scene = host.Scene(name, samples=1, bounces=4)
        scene.update_nodes()
        osc = O.OracleScene(scene)
        d = scene.desc()
        nodes, n_nodes = scene.nodes()
        node_bytes = C.cast(nodes, C.c_void_p)
        models = np.ctypeslib.as_array(C.cast(d.offsets, C.POINTER(C.c_uint32)), shape=(d.offset_count, 16))
        for trial in range(1500):
            ni = int(rng.integers(0, n_nodes))
            model_id = nodes[ni].modelId
            tri_count = int(models[model_id, 1]) // 3
            if tri_count == 0:
                continue
            prim = int(rng.integers(0, tri_count))
            ro = f3(*rng.uniform(-8, 8, 3)); rd_ = rng.normal(size=3); rd = f3(*(rd_ / np.linalg.norm(rd_)))
            a, b = f18(), f18()
            ma = ref.ref_get_material_data(ni + 1, prim, ro, rd, node_bytes, C.cast(d.offsets, C.c_void_p), C.cast(d.vertices, C.c_void_p), C.cast(d.indices, C.c_void_p), a)
            mb = L.gko_get_material_data(osc._h, ni + 1, prim, ro, rd, b)
            assert ma == mb
            ba, bb = bits(a), bits(b)
            keep = np.r_[0:6, 9:18]                      # Position, Normal, TexCoord, LocalNormal, LocalTangent
            both_nan = np.isnan(np.array(a)) & np.isnan(np.array(b))  # degenerate ray / triangle pairs: 1/0 -> inf/NaN identically
            assert ((ba == bb) | both_nan)[keep].all(), (name, ni, prim, list(a), list(b))
            ta, tb = np.array(a)[6:9], np.array(b)[6:9]
            assert both_nan[6:9].all() or np.abs(ta - tb).max() <= 2e-7
            for off in (0, 3, 15, 16, 40):
                node_ptr = C.c_void_p(node_bytes.value + ni * 208)
                assert ref.ref_fetch_material_id(node_ptr, off) == L.gko_fetch_material_id(osc._h, ni, off)
    for trial in range(3000):
        vnt, n, t = f3(*rng.uniform(0, 1, 3)), f3(*rng.normal(size=3)), f4(*rng.normal(size=3), float(rng.choice([0.0, 2.0])))
        a, b = f3(), f3()
        ref.ref_mikktspace(vnt, n, t, a); L.gko_mikktspace(vnt, n, t, b)
        assert (bits(a) == bits(b)).all()
        nn, cn = rng.normal(size=3), rng.normal(size=3)
        nn /= np.linalg.norm(nn)
        cn = nn + cn * float(rng.choice([0.0, 0.02, 0.2, 1.0]))
        cn /= np.linalg.norm(cn)
        dist = float(rng.uniform(0, 8))
        for same in (0, 1):
            for centre in (0, 1):
                wa = ref.ref_calculate_weight(dist, same, centre, f3(*nn), f3(*cn))
                wb = L.gko_calculate_weight(dist, same, centre, f3(*nn), f3(*cn))
                assert bits([wa])[0] == bits([wb])[0]


REF_ST = os.path.join(os.path.dirname(O.TINYBVH), "libref_stencils.so")


@pytest.mark.skipif(not os.path.exists(REF_ST), reason="oracle/_ref not built (reference not mounted)")
@pytest.mark.parametrize("W,H", [(67, 45), (64, 40)])
def test_reproject_and_denoise_shaders_are_pinned_to_the_references_source(W, H):
    """the reference's Process.ReProject.comp.slang and Process.DenoiseJBF.comp.slang compiled as C++ from the files
    where they lie (oracle/ref_stencils.cpp: storage images in the reference's formats, an 8 x 8 thread group with a real barrier for the
    groupshared tile) and run on random storage images, against the oracle's ReProjectPixel / DenoisePixel. ACCUM_* must be bit-identical in
    every mode (temporal with 16 and 1 frames, progressive, fast reproject); RT_DENOISED bit-identical as well, SDR and HDR - both sides use
    this machine's libm for exp / pow."""
    import ctypes as C
    from gknextengine_b200 import RT, GkcuFrameParams, host
    ref = C.CDLL(REF_ST)
    ref.ref_reproject.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    ref.ref_denoise.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_void_p]
    rng = np.random.default_rng(23)

    def rnd16(hi=4.0):
        return rng.uniform(0, hi, (H, W, 4)).astype(np.float16)

    ids0 = rng.integers(0, 4, (H, W)).astype(np.uint32)
    ids0[rng.uniform(size=(H, W)) < 0.05] = 65535
    ids1 = np.where(rng.uniform(size=(H, W)) < 0.8, ids0, rng.integers(0, 4, (H, W))).astype(np.uint32)
    nrm = rng.normal(size=(H, W, 4)).astype(np.float32)
    nrm[..., :3] /= np.linalg.norm(nrm[..., :3], axis=2, keepdims=True)
    nrm[::2] = np.float32([0, 1, 0, 0.3])
    light = rnd16(); light[5:9, 7:12, :3] = 2000.0   # a light source: the bilateral weights underflow around it (NaN pixels must match too)
    images = {
        "SINGLE_DIFFUSE": light, "SINGLE_SPECULAR": rnd16(), "ALBEDO": rnd16(1.0), "NORMAL": nrm.astype(np.float16),
        "SINGLE_PREV_DIFFUSE": rnd16(), "SINGLE_PREV_SPECULAR": rnd16(), "SINGLE_PREV_ALBEDO": rnd16(1.0),
        "OBJEDCTID_0": ids0, "OBJEDCTID_1": ids1,
        "MOTIONVECTOR": np.where(rng.uniform(size=(H, W, 1)) < 0.5, 0, rng.uniform(-6, 6, (H, W, 2))).astype(np.float32),
        "MOTIONMOMENT": (rng.uniform(size=(H, W)) < 0.1).astype(np.uint16) * 3,
    }
    ofr = O.OracleFrame(W, H)
    for mode in (dict(progressive=0, temporal_frames=16), dict(progressive=0, temporal_frames=1), dict(progressive=1, temporal_frames=16),
                 dict(progressive=0, temporal_frames=8, fast_reproject=1)):
        for hdr in (False, True):
            scene = host.Scene("CornellBox.proc", denoiser=True, hdr=hdr)
            ubo = scene.ubo(W, H, 1)
            p = GkcuFrameParams(spp_world=1, **mode)
            # reference side: its own copies of every image, ACCUM_* seeded with PREV_* (the state the copy pass leaves behind)
            mine = {k: v.copy() for k, v in images.items()}
            for acc, prev in (("ACCUMLATE_DIFFUSE", "SINGLE_PREV_DIFFUSE"), ("ACCUMLATE_SPECULAR", "SINGLE_PREV_SPECULAR"), ("ACCUMLATE_ALBEDO", "SINGLE_PREV_ALBEDO")):
                mine[acc] = images[prev].copy()
            mine["DENOISED"] = np.zeros((H, W, 4), np.float16)
            table = (C.c_void_p * 32)()
            for k, v in mine.items():
                table[RT[k]] = v.ctypes.data
            ref.ref_reproject(table, W, H, mode.get("progressive", 0), mode["temporal_frames"], mode.get("fast_reproject", 0))
            ref.ref_denoise(table, W, H, C.cast(C.byref(ubo), C.c_void_p))
            # oracle side
            for k, v in images.items():
                ofr.image(RT[k])[...] = v
            for acc, prev in (("ACCUMLATE_DIFFUSE", "SINGLE_PREV_DIFFUSE"), ("ACCUMLATE_SPECULAR", "SINGLE_PREV_SPECULAR"), ("ACCUMLATE_ALBEDO", "SINGLE_PREV_ALBEDO")):
                ofr.image(RT[acc])[...] = images[prev]
            ofr.reproject(p)
            ofr.denoise(ubo)
            for name in ("ACCUMLATE_DIFFUSE", "ACCUMLATE_SPECULAR", "ACCUMLATE_ALBEDO", "DENOISED"):
                a, b = mine[name].view(np.uint16), ofr.image(RT[name]).view(np.uint16)
                bad = (a != b).any(axis=2)
                assert not bad.any(), (name, mode, hdr, int(bad.sum()), np.argwhere(bad)[:3].tolist())
    assert np.isnan(mine["DENOISED"].astype(np.float32)).any()  # the light's neighbourhood: 0 / 0 in the reference's arithmetic, reproduced


REF_PT = os.path.join(os.path.dirname(O.TINYBVH), "libref_pathtrace.so")


def _ref_pathtrace_lib():
    import ctypes as C
    ref = C.CDLL(REF_PT)
    ref.ref_pathtrace.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_int, C.POINTER(C.c_uint64)]
    ref.ref_interpolate_ambient_cubes.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float)]
    return ref


def _random_probe_grid(osc, scene, rng):
    """Real voxel distances (the oracle's voxeliser, pinned through the GPU parity tests and the CPU bake tests) under random packed colours."""
    osc.set_lights(scene)
    osc.probes_voxelize()
    cubes, vox, _ = osc.probes_read()
    words = cubes.view(np.uint32).reshape(len(cubes), -1)
    words[...] = rng.integers(0, 2**32, words.shape, dtype=np.uint64).astype(np.uint32)
    osc.set_probes(cubes, vox)
    return cubes, vox


@pytest.mark.skipif(not os.path.exists(REF_PT), reason="oracle/_ref not built (reference not mounted)")
def test_probe_lookup_is_pinned_to_the_references_source():
    """interpolateAmbientCubesV2<Full | DI> (AmbientCube.slang:290-382) compiled as C++ from the file where it lies, against
    the oracle's interpolateAmbientCubesT on the voxelised Cornell box with random probe colours: points inside and outside the grid, on
    cell boundaries, next to walls (occluded probes) and in free space."""
    import ctypes as C
    ref = _ref_pathtrace_lib()
    scene = host.Scene("CornellBox.proc", samples=1, bounces=4)
    scene.update_nodes()
    osc = O.OracleScene(scene)
    rng = np.random.default_rng(3)
    cubes, vox = _random_probe_grid(osc, scene, rng)
    L = osc.L
    L.gko_interpolate_ambient_cubes.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float)]
    pts = np.concatenate([rng.uniform((-3.2, -0.5, -3.2), (3.2, 6.0, 3.2), (4000, 3)), rng.uniform(-30, 30, (300, 3)),
                          np.round(rng.uniform((-3, 0, -3), (3, 5.5, 3), (300, 3)) * 4) / 4 + np.float32([0, -0.375 + 0.03125, 0])]).astype(np.float32)
    nrm = rng.normal(size=(len(pts), 3)).astype(np.float32)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    nrm[::7] = np.float32([0, 1, 0])
    a, b = (C.c_float * 4)(), (C.c_float * 4)()
    nonzero = 0
    for which in (0, 2):
        for p, n in zip(pts, nrm):
            pp, nn = (C.c_float * 3)(*p), (C.c_float * 3)(*n)
            ref.ref_interpolate_ambient_cubes(which, cubes.ctypes.data, vox.ctypes.data, pp, nn, a)
            L.gko_interpolate_ambient_cubes(osc._h, which, pp, nn, b)
            assert np.array_equal(np.float32(list(a)).view(np.uint32), np.float32(list(b)).view(np.uint32)), (which, p, n, list(a), list(b))
            nonzero += any(a[:3])
    assert nonzero > 2000


@pytest.mark.skipif(not os.path.exists(REF_PT), reason="oracle/_ref not built (reference not mounted)")
@pytest.mark.parametrize("name,setup", [("CornellBox.proc", "textures"), ("RTIO.proc", "sky+sun+aperture"), ("CornellBox.proc", "fastgather")])
def test_rt_pass_is_pinned_to_the_references_source(name, setup):
    """the reference's Core.PathTracing.comp.slang and Util.BufferClear.comp.slang compiled whole as C++, with
    the classes of Shading.slang they run (FVisibilityBufferRayCasterV2, FHardwareRayTracerV2, FHardwareDirectIlluminatorV2,
    FPathTracingRendererV2, CalculateMotionVectorV2, FetchGBufferV2, SampleIBLV2) and interpolateAmbientCubesV2, against the oracle's RT pass.
    The ray query, the texture sampler and the visibility buffer are the oracle's (they are hardware in the reference, not shader text; see
    oracle/ref_pathtrace.cpp), so what is pinned is the integrator: control flow, RNG order, shading arithmetic and every image store. Every
    storage image the pass writes must be bit-identical over three frames with moving geometry, and so must the ray counts."""
    import ctypes as C
    ref = _ref_pathtrace_lib()
    W, H = 52, 30  # not a multiple of the 8 x 8 group: out-of-range threads run and their stores are dropped
    rng = np.random.default_rng(5)
    scene = host.Scene(name, samples=3, bounces=4, taa=True, fast_gather=(setup == "fastgather"))
    sky = None
    if "sky" in setup:
        scene.env(has_sky=True, has_sun=True)
        sky = host.make_sky(64, 32)
    scene.update_nodes()
    osc, ofr = O.OracleScene(scene, sky), O.OracleFrame(W, H)
    _random_probe_grid(osc, scene, rng)
    if setup == "textures":
        nrm = rng.normal(size=(32, 32, 3)) * 0.35 + np.array([0, 0, 1.0])
        nrm /= np.linalg.norm(nrm, axis=2, keepdims=True)
        for idx, tex in ((1, rng.uniform(0.2, 1.0, (16, 16, 4))), (2, rng.uniform(0.0, 1.0, (8, 32, 4))), (3, np.concatenate([nrm * 0.5 + 0.5, np.ones((32, 32, 1))], axis=2))):
            osc.set_sky(idx, tex.astype(np.float32))
        mats, nm = scene.materials()
        for i in range(nm):
            if mats[i].MaterialModel != 4:
                mats[i].DiffuseTextureId, mats[i].MRATextureId, mats[i].NormalTextureId = 1, 2, 3
                mats[i].Fuzziness = 0.8
                mats[i].Metalness = 0.5 if i % 2 else 0.0
        osc.L.gko_update_materials(osc._h, mats, nm)
    osc.L.gko_cast_visibility.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
    mine = {"SINGLE_DIFFUSE": np.full((H, W, 4), 7, np.float16), "SINGLE_SPECULAR": np.full((H, W, 4), 7, np.float16), "ALBEDO": np.zeros((H, W, 4), np.float16),
            "NORMAL": np.zeros((H, W, 4), np.float16), "OBJEDCTID_0": np.zeros((H, W), np.uint32), "MOTIONVECTOR": np.zeros((H, W, 2), np.float32),
            "PREV_DEPTHBUFFER": np.zeros((H, W), np.float32), "MOTIONMOMENT": np.zeros((H, W), np.uint16), "MINIGBUFFER": np.zeros((H, W), np.uint32)}
    table = (C.c_void_p * 64)()  # [32 + rt] = optional fp32 shadow images (diagnostics), unused here
    for k, v in mine.items():
        table[RT[k]] = v.ctypes.data
    p = GkcuFrameParams(temporal_frames=16, spp_world=1)
    counters = (C.c_uint64 * 2)()
    moved = False
    for f in range(1, 4):
        if f == 2:  # move the scene between frames: motion vectors and the motion-moment counter (4 x samples on moving pixels)
            scene.jitter_instances(f, 0.6)
            scene.update_nodes()
            osc.update_nodes(scene)
        ubo = scene.ubo(W, H, f, model_rotation_deg=4.0 * f)
        if "aperture" in setup:
            ubo.Aperture, ubo.FocusDistance = 0.4, 6.0
        osc.L.gko_cast_visibility(osc._h, C.byref(ubo), W, H, mine["MINIGBUFFER"].ctypes.data)
        ref.ref_pathtrace(osc._h, C.byref(ubo), table, W, H, counters)
        ofr.render(osc, ubo, p, run_post=False)
        for k in ("OBJEDCTID_0", "MOTIONMOMENT", "ALBEDO", "NORMAL", "MOTIONVECTOR", "PREV_DEPTHBUFFER", "SINGLE_DIFFUSE", "SINGLE_SPECULAR"):
            a, b = mine[k], ofr.image(RT[k])
            ua, ub = a.view(np.uint16 if a.dtype.itemsize == 2 else np.uint32), b.view(np.uint16 if b.dtype.itemsize == 2 else np.uint32)
            bad = (ua != ub).reshape(H, W, -1).any(axis=2)
            assert not bad.any(), (k, f, int(bad.sum()), np.argwhere(bad)[:3].tolist(), a[bad][:2].tolist(), b[bad][:2].tolist())
        c = ofr.counters()
        assert (counters[0], counters[1]) == (c["closest"], c["any"]), (f, list(counters), c)
        moved |= bool(mine["MOTIONMOMENT"].any())
        assert counters[0] > W * H and (counters[1] > W * H // 4) == ("sun" in setup)  # sun shadow rays only where there is a sun
        assert np.isfinite(mine["SINGLE_DIFFUSE"].astype(np.float32)).all() and mine["SINGLE_DIFFUSE"].astype(np.float32).max() > 0.05
    assert moved and (mine["OBJEDCTID_0"] != 65535).mean() > 0.3


@pytest.mark.skipif(not os.path.exists(REF_PT), reason="oracle/_ref not built (reference not mounted)")
@pytest.mark.parametrize("name,sky", [("CornellBox.proc", False), ("RTIO.proc", True)])
def test_probe_bake_is_pinned_to_the_references_source(name, sky):
    """the reference's Bake.HwAmbientCube.comp.slang compiled whole as C++ with FGpuProbeGenerator (RenderAmbientCubeOnly,
    FaceTask, FetchLight), EvaluateSH / SampleIBLRough and the DI probe lookup, against the oracle's BakeCube on probe rows through the scene.
    The reference's threads read neighbouring probes while other threads update them (a race on the GPU); the oracle and the CUDA kernel read
    a snapshot of the range being baked instead. To compare the two deterministically the probes of the baked rows are flagged 'inside
    geometry' (second distance byte 0), which every lookup skips (AmbientCube.slang:327) while the bake itself only looks at the first byte -
    so all light comes from probes outside the rows and the order of the threads cannot matter. Cubes (14 words), ages and ray counts must be
    bit-identical after two bake passes (the second one blends into the first: LerpPackedColorAlt, sky visibility bytes, jitter by age)."""
    import ctypes as C
    from gknextengine_b200._abi import GK_CUBE_SIZE_XY
    ref = _ref_pathtrace_lib()
    ref.ref_probes_bake.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint64)]
    rng = np.random.default_rng(9)
    scene = host.Scene(name, samples=1, bounces=4)
    skyimg = None
    if sky:
        scene.env(has_sky=True, has_sun=True)
        skyimg = host.make_sky(64, 32)
    scene.update_nodes()
    osc = O.OracleScene(scene, skyimg)
    # the reference evaluates the SH sky on every miss, enabled or not (AmbientCube.slang:515), so a set of coefficients must be bound
    osc.set_sky_sh(0, host.project_sh(skyimg if sky else np.full((16, 32, 4), 0.25, np.float32)))
    cubes, vox = _random_probe_grid(osc, scene, rng)
    ubo = scene.ubo(64, 36, 1)
    # rows of 64 probes along x through the middle of the scene at several heights / depths
    rows = []
    for (y, z) in ((3, 96), (9, 90), (14, 101), (22, 96), (2, 106)):
        first = y * GK_CUBE_SIZE_XY * GK_CUBE_SIZE_XY + z * GK_CUBE_SIZE_XY + 64
        rows.append(first)
        vox["distanceToSolid_gg_z01"][first:first + 64] &= np.uint32(0xFFFF00FF)  # inside byte = 0: never looked up
    osc.set_probes(cubes, vox)
    mine_c, mine_v = cubes.copy(), vox.copy()
    counters = (C.c_uint64 * 2)()
    baked = 0
    for rep in range(2):
        for first in rows:
            ref.ref_probes_bake(osc._h, C.byref(ubo), mine_c.ctypes.data, mine_v.ctypes.data, first, 64, counters)
            osc.probes_bake(ubo, first, 64)
            oc, ov, _ = osc.probes_read()
            a, b = mine_c[first:first + 64].view(np.uint32).reshape(64, -1), oc[first:first + 64].view(np.uint32).reshape(64, -1)
            assert np.array_equal(a, b), (rep, first, np.argwhere(a != b)[:4].tolist())
            assert np.array_equal(mine_v["age"][first:first + 64], ov["age"][first:first + 64])
            n_baked = int((mine_v["age"][first:first + 64] == rep + 1).sum())
            baked += n_baked
            assert counters[0] == n_baked * 6 * 16, (list(counters), n_baked)
    # untouched outside the rows
    mask = np.ones(len(cubes), bool)
    for first in rows:
        mask[first:first + 64] = False
    assert np.array_equal(mine_c[mask].view(np.uint32), cubes[mask].view(np.uint32))
    assert baked > 100


REF_VX = os.path.join(os.path.dirname(O.TINYBVH), "libref_voxelizer.so")


@pytest.mark.skipif(not os.path.exists(REF_VX), reason="oracle/_ref not built (reference not mounted)")
@pytest.mark.parametrize("name", ["CornellBox.proc", "RTIO.proc"])
def test_voxeliser_and_page_index_are_pinned_to_the_references_source(name):
    """the reference's CPU voxeliser (VoxelizeCube, InsideGeometry, DetectDistance, PackBytes) and page index
    (FCPUPageIndex::UpdateData / GetPage), line ranges of src/Assets/CPUAccelerationStructure.cpp compiled unmodified (oracle/ref_voxelizer.cpp),
    against the oracle's VoxelizeCube / UpdatePageIndex: the four VoxelData words of every probe in slabs through the scene and all 4096 page
    records must be identical."""
    import ctypes as C
    from gknextengine_b200._abi import GK_CUBE_SIZE_XY
    from gknextengine_b200.renderer import CudaPathTracingRenderer as R
    ref = C.CDLL(REF_VX)
    ref.ref_voxelize.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
    ref.ref_page_index.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
    scene = host.Scene(name, samples=1, bounces=4)
    scene.update_nodes()
    osc = O.OracleScene(scene)
    osc.set_lights(scene)
    osc.probes_voxelize()
    _, vox, pages = osc.probes_read(pages=True)
    layer = GK_CUBE_SIZE_XY * GK_CUBE_SIZE_XY
    checked = solid = 0
    for y in (0, 1, 2, 5, 11, 23, 30):   # horizontal slabs of the grid, 13 824 probes each
        first, count = y * layer + 60 * GK_CUBE_SIZE_XY, 72 * GK_CUBE_SIZE_XY   # rows z = 60 .. 131 around the origin
        mine = np.zeros(count, dtype=R.VOXEL_DTYPE)
        ref.ref_voxelize(osc._h, first, count, mine.ctypes.data)
        theirs = vox[first:first + count]
        for f in ("matId", "age", "distanceToSolid_gg_z01", "distanceToSolid_x01_y01"):
            bad = mine[f] != theirs[f]
            assert not bad.any(), (y, f, int(bad.sum()), np.argwhere(bad)[:3].tolist(), mine[f][bad][:3].tolist(), theirs[f][bad][:3].tolist())
        checked += count
        solid += int((mine["matId"] != 0).sum())
    assert checked > 90000 and solid > 50
    mine_pages = np.zeros(len(pages), dtype=R.PAGE_DTYPE)
    ref.ref_page_index(vox.ctypes.data, len(vox), mine_pages.ctypes.data)
    assert np.array_equal(mine_pages.view(np.uint32), pages.view(np.uint32)) and int(mine_pages["voxelCount"].sum()) == int((vox["matId"] != 0).sum()) > 0



@pytest.mark.skipif(not os.path.exists(REF_PT) or not os.path.exists(REF_ST), reason="oracle/_ref not built (reference not mounted)")
def test_whole_frames_over_time_are_pinned_to_the_references_source():
    """The passes together, frame after frame: BufferClear -> PathTracing -> ReProject -> DenoiseJBF -> copy pass, all from the reference's
    compiled shader text (helpers.CompiledReferenceFrame), against the oracle's gko_render_frame + gko_end_frame_copies over six frames of the
    textured Cornell box with geometry that moves in frame 3 (history rejection by the motion moment, 4 x samples on moving pixels, temporal
    blend afterwards). Every storage image bit-identical, RT_DENOISED included (same libm on both sides)."""
    from helpers import CompiledReferenceFrame, textured_cornell_materials
    W, H = 72, 44
    rng = np.random.default_rng(17)
    scene = host.Scene("CornellBox.proc", samples=2, bounces=4, taa=True, denoiser=True)
    scene.update_nodes()
    osc, ofr = O.OracleScene(scene), O.OracleFrame(W, H)
    textures, mats, nm = textured_cornell_materials(scene, rng)
    for idx, tex in textures:
        osc.set_sky(idx, tex)
    osc.L.gko_update_materials(osc._h, mats, nm)
    _random_probe_grid(osc, scene, rng)
    ref = CompiledReferenceFrame(W, H)
    p = GkcuFrameParams(temporal_frames=16, spp_world=1)
    for f in range(1, 7):
        if f == 3:
            scene.jitter_instances(2, 0.6)
            scene.update_nodes()
            osc.update_nodes(scene)
        ubo = scene.ubo(W, H, f)
        ref.render(osc, ubo)
        ofr.render(osc, ubo, p)
        for name in ("OBJEDCTID_0", "ALBEDO", "NORMAL", "MOTIONVECTOR", "MOTIONMOMENT", "PREV_DEPTHBUFFER", "SINGLE_DIFFUSE", "SINGLE_SPECULAR",
                     "ACCUMLATE_DIFFUSE", "ACCUMLATE_SPECULAR", "ACCUMLATE_ALBEDO", "DENOISED"):
            bad = ref.mismatches(name, ofr.image(RT[name]))
            assert not bad.any(), (name, f, int(bad.sum()), np.argwhere(bad)[:3].tolist())
        c = ofr.counters()
        assert (ref.closest, ref.any) == (c["closest"], c["any"])
        ref.end_frame()
        ofr.end_frame_copies()
    assert ref.images["MOTIONMOMENT"].any() and ref.images["ACCUMLATE_DIFFUSE"].astype(np.float32).max() > 0.05


@pytest.mark.skipif(not os.path.exists(REF_SH), reason="oracle/_ref not built (reference not mounted)")
def test_sky_sh_projection_is_pinned_to_the_references_source():
    """ProjectHdrToSh (src/Assets/Texture.cpp:180-258), the host-side SH9 projection of the sky that feeds the probe bake, compiled
    unmodified from the file where it lies, against the host library's restatement (gkhost_project_sh): all 27 coefficients bit-identical on
    the procedural sky and on random images of odd sizes (same libm on both sides)."""
    import ctypes as C
    ref = C.CDLL(REF_SH)
    ref.ref_project_sh.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_float * 27)]
    rng = np.random.default_rng(12)
    for img in (host.make_sky(64, 32), host.make_sky(256, 128), rng.uniform(0, 30, (17, 41, 4)).astype(np.float32), rng.uniform(0, 2, (1, 1, 4)).astype(np.float32)):
        img = np.ascontiguousarray(img, dtype=np.float32)
        out = (C.c_float * 27)()
        ref.ref_project_sh(img.ctypes.data, img.shape[1], img.shape[0], out)
        mine = host.project_sh(img)
        got = np.float32([[mine.coefficients[c][i] for i in range(9)] for c in range(3)]).reshape(-1)
        assert np.array_equal(got.view(np.uint32), np.float32(list(out)).view(np.uint32)), (img.shape, got[:4], list(out)[:4])
        assert np.abs(got).max() > 0


REF_PM = os.path.join(os.path.dirname(O.TINYBVH), "libref_procmodel.so")


@pytest.mark.skipif(not os.path.exists(REF_PM), reason="oracle/_ref not built (reference not mounted)")
def test_procedural_models_and_materials_are_pinned_to_the_references_source():
    """The inputs of BASELINE configs 1 and 2: FProcModel::CreateCornellBox / CreateBox / CreateSphere (src/Assets/FProcModel.cpp:9-254) and the
    Material constructors (src/Assets/Material.hpp:8-81), compiled unmodified from the files where they lie (oracle/ref_procmodel.cpp), against
    the host library's generators (gkhost_proc_model / gkhost_material): every vertex (position, normal, tangent, uv, material slot), every
    index, the four Cornell materials, the quad light and all five material kinds bit-identical."""
    import ctypes as C
    from gknextengine_b200.host import lib as hostlib
    ref, H = C.CDLL(REF_PM), hostlib()
    sig = [C.c_int, C.POINTER(C.c_float), C.c_void_p, C.POINTER(C.c_uint32), C.c_void_p, C.POINTER(C.c_uint32), C.c_void_p, C.c_void_p]
    for fn in (ref.ref_proc_model, H.gkhost_proc_model):
        fn.argtypes, fn.restype = sig, C.c_uint32
    for fn in (ref.ref_material, H.gkhost_material):
        fn.argtypes = [C.c_int, C.POINTER(C.c_float), C.c_float, C.c_float, C.c_void_p]

    def gen(fn, kind, args):
        a = (C.c_float * 6)(*args)
        nv, ni = C.c_uint32(), C.c_uint32()
        fn(kind, a, None, C.byref(nv), None, C.byref(ni), None, None)
        v, i = np.zeros((nv.value, 13), np.uint32), np.zeros(ni.value, np.uint32)
        mats, light = np.zeros((8, 16), np.uint32), np.zeros(20, np.uint32)
        nm = fn(kind, a, v.ctypes.data, C.byref(nv), i.ctypes.data, C.byref(ni), mats.ctypes.data, light.ctypes.data)
        return v, i, mats[:nm, :14], light  # 14 words: the 56 bytes the reference's struct defines (the rest is padding there, reserved words here)

    rng = np.random.default_rng(8)
    cases = [(0, [5.55]), (0, [1.0]), (1, [-1000, -0.5, -1000, 1000, 0, 1000]), (1, [-0.8, 0, -0.8, 0.8, 1.6, 0.8]), (2, [0, 0, 0, 0.2]), (2, [0, 0, 0, 1.0]),
             (2, [*rng.uniform(-5, 5, 3), 0.37])]
    for kind, args in cases:
        args = list(args) + [0.0] * (6 - len(args))
        rv, ri, rm, rl = gen(ref.ref_proc_model, kind, args)
        hv, hi, hm, hl = gen(H.gkhost_proc_model, kind, args)
        assert rv.shape == hv.shape and np.array_equal(rv, hv), (kind, args, np.argwhere(rv != hv)[:4].tolist())
        assert np.array_equal(ri, hi) and np.array_equal(rm, hm) and np.array_equal(rl, hl), (kind, args)
    assert gen(ref.ref_proc_model, 2, [0, 0, 0, 0.2, 0, 0])[0].shape == (561, 13) and len(gen(ref.ref_proc_model, 0, [5.55] + [0] * 5)[1]) == 36
    d = (C.c_float * 3)(0.3, 0.6, 0.9)
    for kind, a, b in ((0, 0, 0), (1, 0.25, 0), (2, 0.5, 0), (3, 1.5, 0.1), (4, 0, 0)):
        x, y = np.zeros(16, np.uint32), np.zeros(16, np.uint32)
        ref.ref_material(kind, d, a, b, x.ctypes.data)
        H.gkhost_material(kind, d, a, b, y.ctypes.data)
        assert np.array_equal(x[:14], y[:14]), (kind, x[:14].tolist(), y[:14].tolist())


@pytest.mark.skipif(not os.path.exists(REF_PM), reason="oracle/_ref not built (reference not mounted)")
@pytest.mark.parametrize("name", ["CornellBox.proc", "RTIO.proc"])
def test_procedural_scenes_are_pinned_to_the_references_source(name):
    """The scenes BASELINE configs 1 and 2 are rendered from: SceneList's CornellBox and RayTracingInOneWeekend (src/Runtime/SceneList.cpp:27-227,
    with LoadScene's root material, :262), compiled unmodified (oracle/ref_procmodel.cpp) against the host library's scene of the same name:
    node for node the translation (bit pattern), model index, instance id and the 16 material slots, and the whole material table - for
    RTIO.proc that is the mt19937(42) stream deciding 1 935 sphere positions and 1 940 materials. Rotations, scales and the camera go through
    glm in the engine (not in the mount) and are not compared here beyond scale = 1 for the RTIOW spheres."""
    import ctypes as C
    from gknextengine_b200._abi import GkMaterial, GkNodeProxy
    ref = C.CDLL(REF_PM)
    ref.ref_scene.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint32 * 4), C.POINTER(C.c_uint32 * 2)]
    counts, flags = (C.c_uint32 * 4)(), (C.c_uint32 * 2)()
    assert ref.ref_scene(name.encode(), None, None, counts, flags) == 1
    nn, nm = counts[0], counts[1]
    rn, rm = np.zeros((nn, 24), np.uint32), np.zeros((nm, 16), np.uint32)
    ref.ref_scene(name.encode(), rn.ctypes.data, rm.ctypes.data, counts, flags)
    scene = host.Scene(name)
    scene.update_nodes()
    hp, hn = scene.nodes()
    mp, hm = scene.materials()
    assert (hn, hm) == (nn, nm) and scene.counts()["models"] == counts[2], (hn, hm, list(counts))
    if name == "RTIO.proc":
        assert (nn, nm) == (1939, 1940)
    hnodes = np.frombuffer((C.c_uint8 * (C.sizeof(GkNodeProxy) * hn)).from_address(C.addressof(hp.contents) if hasattr(hp, "contents") else C.cast(hp, C.c_void_p).value),
                           dtype=np.uint32).reshape(hn, -1)
    # GkNodeProxy words: instanceId 0, modelId 1, visible 2, nort 3, worldTS 4..19 (column-major: translation = words 16..18), prevTS 20..35, matId 36..51
    assert np.array_equal(hnodes[:, 16:19], rn[:, 0:3]), np.argwhere(hnodes[:, 16:19] != rn[:, 0:3])[:4].tolist()
    # NodeProxy.modelId = model * 10 + section (Scene.cpp:475); these models have one section
    assert np.array_equal(hnodes[:, 1], rn[:, 6] * 10) and np.array_equal(hnodes[:, 0], rn[:, 7]) and np.array_equal(hnodes[:, 36:52], rn[:, 8:24])
    if name == "RTIO.proc":
        ident = np.float32([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0]).view(np.uint32)
        assert (hnodes[:, 4:16] == ident).all() and (rn[:, 3:6] == np.float32(1).view(np.uint32)).all()
    hmats = np.frombuffer((C.c_uint8 * (C.sizeof(GkMaterial) * hm)).from_address(C.cast(mp, C.c_void_p).value), dtype=np.uint32).reshape(hm, 16)
    assert np.array_equal(hmats[:, :14], rm[:, :14]), np.argwhere(hmats[:, :14] != rm[:, :14])[:4].tolist()
    ubo = scene.ubo(64, 36, 1)
    assert (bool(ubo.HasSky), bool(ubo.HasSun)) == (bool(flags[0]), bool(flags[1]))


@pytest.mark.skipif(not os.path.exists(REF_PM), reason="oracle/_ref not built (reference not mounted)")
def test_taa_jitter_is_pinned_to_the_references_halton_sequence():
    """The sub-pixel jitter of the projection (Engine.cpp:771-777) comes from HaltonSequence (Engine.cpp:642-651, compiled unmodified): the host
    library's UBO must carry (halton(i + 1, 2) - 0.5) / width * 2 and (halton(i + 1, 3) - 0.5) / height * 2 in Projection[2][0..1], bit for bit,
    for every frame of the 16-frame cycle and beyond."""
    import ctypes as C
    ref = C.CDLL(REF_PM)
    ref.ref_halton.restype, ref.ref_halton.argtypes = C.c_float, [C.c_int, C.c_int]
    W, H, TF = 1920, 1080, 16
    scene = host.Scene("CornellBox.proc", taa=True, temporal_frames=TF)
    f = np.float32
    for frame in list(range(0, 40)) + [1000, 65535]:
        u = scene.ubo(W, H, frame)
        i = frame % TF
        jx = (f(ref.ref_halton(i + 1, 2)) - f(0.5)) / f(W) * f(2.0)
        jy = (f(ref.ref_halton(i + 1, 3)) - f(0.5)) / f(H) * f(2.0)
        got = np.float32([u.Projection.m[2][0], u.Projection.m[2][1]])
        assert np.array_equal(got.view(np.uint32), np.float32([jx, jy]).view(np.uint32)), (frame, got, jx, jy)
    u = host.Scene("CornellBox.proc", taa=False).ubo(W, H, 3)
    assert u.Projection.m[2][0] == 0.0 and u.Projection.m[2][1] == 0.0
