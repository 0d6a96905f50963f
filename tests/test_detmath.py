"""include/gk_detmath.h: the pinned sin/cos used for ray directions stays within 2 ulp of libm on [0, 2*pi]."""
import ctypes as C

import numpy as np

import oracle_lib as O


def test_det_sincos_accuracy():
    L = O.load()
    xs = np.concatenate([np.linspace(0, 2 * np.pi, 20001, dtype=np.float32), np.float32([0, np.pi / 4, np.pi / 2, np.pi, 1.5 * np.pi, 2 * np.pi])])
    s, c = C.c_float(), C.c_float()
    worst = 0.0
    for x in xs:
        L.gko_det_sincos(float(x), C.byref(s), C.byref(c))
        for got, ref in ((s.value, np.sin(np.float64(x))), (c.value, np.cos(np.float64(x)))):
            ulp = np.spacing(np.float32(max(abs(ref), 2.0 ** -20)))
            worst = max(worst, abs(got - ref) / ulp)
    assert worst < 2.0, worst


def test_det_sincos_negative_and_symmetry():
    L = O.load()
    s, c, s2, c2 = C.c_float(), C.c_float(), C.c_float(), C.c_float()
    for x in (0.1, 0.7, 2.0, 5.5):
        L.gko_det_sincos(x, C.byref(s), C.byref(c))
        L.gko_det_sincos(-x, C.byref(s2), C.byref(c2))
        assert s2.value == -s.value and c2.value == c.value
        assert abs(s.value ** 2 + c.value ** 2 - 1) < 3e-7


def test_constant_division_by_multiply_and_two_fmas_is_exact():
    """gk_shading.cuh div_const: q = a * rb; q + fma(-q, b, a) * rb must equal the IEEE quotient a / b for every operand the probe lookup
    feeds it (10-bit colour codes over 1023, bytes over 255), otherwise the CUDA lookup would drift from the oracle's plain divisions."""
    f = np.float32
    for b, n in ((1023, 1024), (255, 256)):
        a = np.arange(n, dtype=np.float32)
        rb = f(1.0) / f(b)
        q = a * rb
        r = (a.astype(np.float64) - q.astype(np.float64) * b).astype(np.float32)           # fma(-q, b, a): exact product, one rounding
        got = (q.astype(np.float64) + r.astype(np.float64) * np.float64(rb)).astype(np.float32)  # fma(r, rb, q)
        assert np.array_equal(got, a / f(b))


def _det(L, fn, a, b=0.0):
    L.gko_det_math.restype = C.c_float
    L.gko_det_math.argtypes = [C.c_int, C.c_float, C.c_float]
    return L.gko_det_math(fn, float(a), float(b))


def _ulps(got, ref):
    return abs(np.float64(got) - ref) / np.spacing(np.float32(max(abs(ref), 2.0 ** -100)))


def test_det_leaf_transcendentals_accuracy_and_special_values():
    """gk_det_atan2f / acosf / expf / logf / powf (sky lookup, bilateral weights, tonemappers): a few ulp from the exact function over the
    ranges the renderer feeds them, and the C library's conventions at the special values."""
    L = O.load()
    rng = np.random.default_rng(4)
    worst = dict(atan2=0.0, acos=0.0, exp=0.0, log=0.0, pow=0.0)
    for _ in range(20000):
        a, b = np.float32(rng.uniform(-1, 1)), np.float32(rng.uniform(-1, 1))
        worst["atan2"] = max(worst["atan2"], _ulps(_det(L, 0, a, b), np.arctan2(np.float64(a), np.float64(b))))
        worst["acos"] = max(worst["acos"], _ulps(_det(L, 1, a), np.arccos(np.float64(a))))
        x = np.float32(-80.0 * rng.uniform() ** 2)
        worst["exp"] = max(worst["exp"], _ulps(_det(L, 2, x), np.exp(np.float64(x))))
        p = np.float32(10.0 ** rng.uniform(-6, 3))
        worst["log"] = max(worst["log"], _ulps(_det(L, 3, p), np.log(np.float64(p))))
        y = np.float32(rng.choice([1.33, 0.1593017578125, 78.84375, 2.2]))
        base = np.float32(rng.uniform(0.05, 1.2)) if y > 50 else p
        ref = np.float64(base) ** np.float64(y)
        if 1e-30 < ref < 1e30:
            # pow = exp(y log x): the error grows with |y ln x| (like a GPU's intrinsic), so it is measured relative to that
            rel = abs(np.float64(_det(L, 4, base, y)) - ref) / ref
            worst["pow"] = max(worst["pow"], rel / max(1.0, abs(np.float64(y) * np.log(np.float64(base)))))
    assert worst["atan2"] <= 4 and worst["acos"] <= 2 and worst["exp"] <= 2 and worst["log"] <= 2 and worst["pow"] < 4e-7, worst
    inf, nan = np.float32(np.inf), np.float32(np.nan)
    assert _det(L, 0, 0.0, -1.0) == np.float32(np.pi) and _det(L, 0, 0.0, 1.0) == 0.0 and _det(L, 0, 1.0, 0.0) == np.float32(np.pi / 2) and _det(L, 0, -1.0, 0.0) == -np.float32(np.pi / 2)
    assert _det(L, 1, 1.0) == 0.0 and _det(L, 1, -1.0) == np.float32(np.pi) and _det(L, 1, 0.0) == np.float32(np.pi / 2) and np.isnan(_det(L, 1, 1.5))
    assert _det(L, 2, 0.0) == 1.0 and _det(L, 2, -200.0) == 0.0 and _det(L, 2, 100.0) == inf and 0 < _det(L, 2, -100.0) < 1e-40  # gradual underflow
    assert _det(L, 3, 1.0) == 0.0 and _det(L, 3, 0.0) == -inf and np.isnan(_det(L, 3, -1.0)) and _det(L, 3, inf) == inf
    assert _det(L, 4, 0.0, 1.33) == 0.0 and _det(L, 4, 5.0, 0.0) == 1.0 and np.isnan(_det(L, 4, -1.0, 1.33)) and _det(L, 4, inf, 2.0) == inf and _det(L, 4, 1.0, 78.0) == 1.0
    for fn in range(5):
        assert np.isnan(_det(L, fn, nan, 1.0))
