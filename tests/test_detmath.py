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
