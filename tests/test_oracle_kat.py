"""Pins the oracle's small pure functions to tests/golden/kat.json (independent numpy restatements, see make_golden.py)
and to closed-form properties of the samplers (Const_Func.slang:227-340)."""
import ctypes as C
import json
import os

import numpy as np

import oracle_lib as O
from gknextengine_b200 import host

KAT = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kat.json")))


def test_pcg4d_sequence():
    L = O.load()
    v = (C.c_uint32 * 4)(3, 5, 7, 0)
    for expect in KAT["pcg4d_seed_3_5_7_0"]:
        L.gko_pcg4d(C.byref(v))
        assert list(v) == expect


def test_uint_to_float_edges():
    L = O.load()
    for x, expect in KAT["uint_to_float"]:
        got = L.gko_uint_to_float(x)
        assert got == np.float32(expect)
        assert 0.0 <= got < 1.0


def test_gt_tonemap_table():
    L = O.load()
    for x, expect in KAT["gt_tonemap"]:
        i = (C.c_float * 3)(x, x, x)
        o = (C.c_float * 3)()
        L.gko_gt_tonemap(C.byref(i), C.byref(o))
        assert abs(o[0] - expect) <= 2e-6 * max(1.0, abs(expect)), (x, o[0], expect)
        assert o[0] == o[1] == o[2]


def test_schlick_table():
    L = O.load()
    for c, ior, expect in KAT["schlick"]:
        assert abs(L.gko_schlick(c, ior) - expect) <= 1e-6


def test_glm_float16_rounding():
    for f, bits in KAT["to_float16"]:
        assert host.lib().gkhost_to_float16(f) == bits, f
    # exhaustively: every finite half round-trips, and a value just above a tie rounds up in magnitude (glm, not IEEE ties-to-even)
    L = O.load()
    for h in range(0, 0x7C00, 37):
        assert host.lib().gkhost_to_float16(L.gko_half_to_float(h)) == h


def test_storage_half_is_round_to_nearest_even():
    L = O.load()
    assert L.gko_float_to_half(1.00048828125) == 0x3C00  # tie -> even
    assert L.gko_float_to_half(1.00146484375) == 0x3C02  # tie -> even (up)
    assert L.gko_float_to_half(70000.0) == 0x7C00
    assert L.gko_float_to_half(-0.0) == 0x8000


def test_hemisphere_sampler_properties():
    L = O.load()
    seed = (C.c_uint32 * 4)(11, 22, 33, 0)
    out = (C.c_float * 3)()
    zs = []
    for _ in range(4000):
        L.gko_random_in_hemisphere1(C.byref(seed), C.byref(out))
        v = np.array(out[:])
        assert abs(np.linalg.norm(v) - 1) < 2e-6 and v[2] >= 0
        zs.append(v[2])
    assert abs(np.mean(zs) - 2.0 / 3.0) < 0.02  # cosine-weighted: E[cos] = 2/3


def test_ggx_sampling_properties():
    L = O.load()
    n = (C.c_float * 3)(0.0, 0.6, 0.8)
    out = (C.c_float * 3)()
    seed = (C.c_uint32 * 4)(1, 2, 3, 0)
    for rough in (0.0, 0.1, 0.5, 1.0):
        cos = []
        for _ in range(500):
            L.gko_ggx_sampling(C.byref(seed), rough, C.byref(n), C.byref(out))
            v = np.array(out[:])
            assert abs(np.linalg.norm(v) - 1) < 1e-5
            cos.append(float(v @ np.array(n[:])))
        if rough == 0.0:
            assert min(cos) > 1 - 1e-6  # mirror: the microfacet normal is the macro normal
        assert min(cos) > -1e-6
    # rougher lobes spread further
    assert np.mean(cos) < 0.95
