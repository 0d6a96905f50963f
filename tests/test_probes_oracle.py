"""CPU tests of the probe-grid restatement (oracle + host helpers): packed colour format, SH sky,
voxelisation and bake invariants on the Cornell box. The GPU parity of the same functions is in test_gpu_parity.py."""
import ctypes as C

import numpy as np

import oracle_lib as O
from gknextengine_b200 import host
from gknextengine_b200._abi import GK_CUBE_COUNT, GK_CUBE_SIZE_XY


def _pack(L, c):
    return L.gko_pack_rgb10a2((C.c_float * 4)(*c))


def _unpack(L, p):
    out = (C.c_float * 4)()
    L.gko_unpack_rgb10a2(p, out)
    return list(out)


def test_rgb10a2_known_answers():
    """AmbientCube.slang:58-78: 10:10:10:2 over [0, MAX_ILLUMINANCE = 512], truncating; alpha is dropped by the unpack."""
    L = O.load()
    assert _pack(L, (0, 0, 0, 0)) == 0
    assert _pack(L, (512, 512, 512, 512)) == 0xFFFFFFFF
    assert _pack(L, (1e9, -5.0, 256.0, 0.0)) == 1023 | (0 << 10) | (511 << 20)
    assert _pack(L, (0.5, 0.0, 0.0, 0.0)) == 0          # 0.5 / 512 * 1023 = 0.999 -> 0 (truncation, not rounding)
    assert _pack(L, (0.51, 0.0, 0.0, 0.0)) == 1
    f = np.float32
    assert _unpack(L, 1023 | (511 << 10) | (1 << 20) | (3 << 30)) == [512.0, f(f(511) / f(1023)) * f(512), f(f(1) / f(1023)) * f(512), 0.0]
    # pack(unpack(p)) is not the identity: the float round trip truncates some codes one step down, which is why an untouched
    # probe face slowly darkens under LerpPackedColorAlt; codes never increase
    codes = np.arange(1024)
    back = np.array([_pack(L, _unpack(L, int(c))) & 0x3FF for c in codes])
    assert (back <= codes).all() and (codes - back).max() <= 1


def test_sh_projection_of_constant_and_linear_skies():
    """Texture.cpp:180-262 + Shading.slang:146-192: a constant sky projects onto band 0 only and evaluates back to itself; a sky that
    brightens towards +Y evaluates brighter for an up-facing normal."""
    h, w = 64, 128
    sky = np.ones((h, w, 4), np.float32) * np.array([2.0, 1.0, 0.5, 1.0], np.float32)
    sh = host.project_sh(sky)
    co = np.array(sh.coefficients)
    assert np.allclose(co[:, 0], np.array([2.0, 1.0, 0.5]) * 0.282095 * 4 * np.pi, rtol=2e-3)
    assert np.abs(co[:, 1:]).max() < 2e-3
    scene = host.Scene("CornellBox.proc", samples=1, bounces=4)
    scene.update_nodes()
    osc = O.OracleScene(scene)
    osc.set_sky_sh(0, sh)
    out = (C.c_float * 4)()
    for d in ((0, 1, 0), (1, 0, 0), (0.6, -0.8, 0)):
        osc.L.gko_sample_ibl_rough(osc._h, 0, (C.c_float * 3)(*d), 0.25, out)
        assert np.allclose(list(out), [2.0, 1.0, 0.5, 1.0], rtol=3e-3)
    osc.L.gko_sample_ibl_rough(osc._h, 7, (C.c_float * 3)(0, 1, 0), 0.0, out)  # no SH uploaded for that sky: black, alpha 1
    assert list(out) == [0.0, 0.0, 0.0, 1.0]
    y = np.cos((np.arange(h) + 0.5) / h * np.pi)[:, None, None]
    grad = np.broadcast_to(np.maximum(y, 0.0), (h, w, 1)) * np.ones((1, 1, 4), np.float32)
    osc.set_sky_sh(1, host.project_sh(np.ascontiguousarray(grad, np.float32)))
    up, down = (C.c_float * 4)(), (C.c_float * 4)()
    osc.L.gko_sample_ibl_rough(osc._h, 1, (C.c_float * 3)(0, 1, 0), 0.0, up)
    osc.L.gko_sample_ibl_rough(osc._h, 1, (C.c_float * 3)(0, -1, 0), 0.0, down)
    assert up[0] > 0.8 and down[0] < 0.1


def _grid_index(x, y, z):
    return y * GK_CUBE_SIZE_XY * GK_CUBE_SIZE_XY + z * GK_CUBE_SIZE_XY + x


def test_voxelize_and_bake_cornell_invariants():
    scene = host.Scene("CornellBox.proc", samples=1, bounces=4)
    scene.update_nodes()
    osc = O.OracleScene(scene)
    osc.set_lights(scene)
    osc.probes_voxelize()
    cubes, vox, pages = osc.probes_read(pages=True)
    assert len(vox) == GK_CUBE_COUNT and (vox["age"] == 0).all() and not cubes.view(np.uint32).any()
    d0 = vox["distanceToSolid_gg_z01"] & 0xFF
    # probe (x, y, z) sits at (x, y, z) * 0.25 + (-24, -0.34375, -24); the Cornell box spans x, z in [-2.775, 2.775], y in [0, 5.55]
    centre = _grid_index(96, 12, 96)       # (0, 2.656, 0): open space inside the room, walls ~2.7 m away, sphere/boxes nearer
    assert 3 <= d0[centre] <= 12 and vox["matId"][centre] == 0
    assert ((vox["distanceToSolid_gg_z01"] >> 8) & 0xFF)[centre] == 255  # "inside" product: all six axis distances >= one cell
    far = _grid_index(5, 40, 5)            # far outside: nothing within 16 m along the 14 probe rays
    assert d0[far] == (int(255 / 0.25) & 0xFF) and vox["distanceToSolid_x01_y01"][far] == 0xFFFFFFFF
    under_floor = _grid_index(96, 1, 96)   # y = -0.094: 9 cm under the floor quad, sees its back face within one cell -> solid
    assert d0[under_floor] == 0 and vox["matId"][under_floor] != 0
    # page index: every voxel with a material lands in exactly one page; the room occupies the 4 pages around the origin
    assert pages["voxelCount"].sum() == (vox["matId"] != 0).sum() > 1000
    assert (pages["voxelCount"] > 0).sum() == 4 and not pages["pageId"].any()

    ubo = scene.ubo(640, 360, 1)
    assert ubo.LightCount == 1 and not ubo.HasSun
    active = np.nonzero(d0 < 8)[0]
    first, count = int(active.min()), int(active.max() - active.min() + 1)
    for _ in range(3):
        osc.probes_bake(ubo, first, count)
    cubes, vox2, _ = osc.probes_read()
    assert (vox2["age"][active] == 3).all() and (vox2["age"][d0 >= 8] == 0).all()
    assert (vox2["matId"] == vox["matId"]).all()
    # the quad light (2000 units, facing down from the ceiling) lights the up-facing side of probes below it directly ...
    below_light = _grid_index(96, 12, 96)
    r = lambda p: int(p) & 0x3FF
    assert r(cubes["PosY_D"][below_light]) > 0 and r(cubes["NegY_D"][below_light]) == 0
    # ... and from the second round on the walls bounce it back: indirect terms appear, tinted by the wall they face
    # (left wall = material slot 1 = green 0.12/0.45/0.15, right wall = slot 0 = red 0.65/0.05/0.05, FProcModel.cpp:26-29)
    assert (cubes["PosY"] != 0).sum() > 1000
    g = lambda p: (int(p) >> 10) & 0x3FF
    near_left = _grid_index(96 - 9, 12, 96)   # x = -2.25, half a metre from the left wall
    near_right = _grid_index(96 + 9, 12, 96)
    assert g(cubes["NegX"][near_left]) > 2 * r(cubes["NegX"][near_left]) > 0
    assert r(cubes["PosX"][near_right]) > 2 * g(cubes["PosX"][near_right]) >= 0 and r(cubes["PosX"][near_right]) > 2
    # deterministic
    osc2 = O.OracleScene(scene)
    osc2.set_lights(scene)
    osc2.probes_voxelize(threads=3)
    for _ in range(3):
        osc2.probes_bake(ubo, first, count, threads=5)
    c2, v2, _ = osc2.probes_read()
    assert (c2 == cubes).all() and (v2 == vox2).all()
