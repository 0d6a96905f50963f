"""Generates tests/golden/kat.json: known answers for the small pure functions of the path (SURVEY 8c "KATs we create").

The reference has no tests and its shaders cannot run here (no slangc / Vulkan), so these values come from independent
restatements in numpy/pure Python of the published formulas the shaders cite (pcg4d: Jarzynski & Olano 2020 as written at
Const_Func.slang:227-233; uint_to_float :235; the GT tonemapper :100-126; Schlick :8-14), NOT from the oracle. The oracle and
the CUDA kernels are then both checked against this file. Run: python tests/golden/make_golden.py
"""
import json
import os

import numpy as np

M32 = 0xFFFFFFFF


def pcg4d(v):
    v = [(x * 1664525 + 1013904223) & M32 for x in v]
    v[0] = (v[0] + v[1] * v[3]) & M32; v[1] = (v[1] + v[2] * v[0]) & M32; v[2] = (v[2] + v[0] * v[1]) & M32; v[3] = (v[3] + v[1] * v[2]) & M32
    v = [x ^ (x >> 16) for x in v]
    v[0] = (v[0] + v[1] * v[3]) & M32; v[1] = (v[1] + v[2] * v[0]) & M32; v[2] = (v[2] + v[0] * v[1]) & M32; v[3] = (v[3] + v[1] * v[2]) & M32
    return v


def uint_to_float(x):
    return float(np.array([0x3F800000 | (x >> 9)], dtype=np.uint32).view(np.float32)[0] - np.float32(1.0))


def gt(x):
    f = np.float32
    x = f(x)
    P, a, m, l, c, b = f(1), f(0.7), f(0.22), f(0.4), f(1.33), f(0)
    e = f(2.71828)
    l0 = (P - m) * l / a
    L_x = m + a * (x - m)
    T_x = m * np.power(x / m, c, dtype=np.float32) + b
    S0 = m + l0
    S1 = m + a * l0
    C2 = a * P / (P - S1)
    S_x = P - (P - S1) * np.power(e, -(C2 * (x - S0) / P), dtype=np.float32)

    def W(x, e0, e1):
        if x <= e0:
            return f(0)
        if x >= e1:
            return f(1)
        t = (x - e0) / (e1 - e0)
        return t * t * (f(3) - f(2) * t)

    def H(x, e0, e1):
        if x <= e0:
            return f(0)
        if x >= e1:
            return f(1)
        return (x - e0) / (e1 - e0)

    w0 = f(1) - W(x, f(0), m)
    w2 = H(x, m + l0, m + l0)
    w1 = f(1) - w0 - w2
    return float(T_x * w0 + L_x * w1 + S_x * w2)


def schlick(cosine, ior):
    f = np.float32
    r0 = (f(1) - f(ior)) / (f(1) + f(ior))
    r0 = r0 * r0
    x = f(1) - f(cosine)
    return float(r0 + (f(1) - r0) * (x * x * x * x * x))


def main():
    out = {}
    seq, v = [], [3, 5, 7, 0]
    for _ in range(6):
        v = pcg4d(v)
        seq.append(list(v))
    out["pcg4d_seed_3_5_7_0"] = seq
    out["uint_to_float"] = [[x, uint_to_float(x)] for x in (0, 1, 511, 512, 0x7FFFFFFF, 0x80000000, 0xFFFFFDFF, 0xFFFFFFFF)]
    out["gt_tonemap"] = [[x, gt(x)] for x in (0.0, 0.01, 0.1, 0.22, 0.3, 0.5, 0.676, 0.9, 1.5, 4.0, 15.0)]
    out["schlick"] = [[c, i, schlick(c, i)] for c in (0.0, 0.25, 0.5, 1.0) for i in (1.0, 1.45, 1.5)]
    # fp16 packing of glm::detail::toFloat16 (round-half-up in magnitude), checked by hand against the algorithm
    out["to_float16"] = [[1.0, 0x3C00], [0.5, 0x3800], [-2.0, 0xC000], [65504.0, 0x7BFF], [1e-8, 0x0000], [1000.0, 0x63D0], [5.55, 0x458D],
                         [2.775, 0x418D], [0.999, 0x3BFE], [1.0009765625, 0x3C01], [1.00048828125, 0x3C01], [1.000244140625, 0x3C00],
                         [6.103515625e-05, 0x0400], [3.0517578125e-05, 0x0200], [70000.0, 0x7C00]]
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "kat.json"), "w") as fjson:
        json.dump(out, fjson, indent=1)


if __name__ == "__main__":
    main()
