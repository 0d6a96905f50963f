"""Shared helpers of the parity tests: build a scene, render it with the CUDA renderer and with the CPU oracle, compare."""
import numpy as np

import oracle_lib as O
from gknextengine_b200 import RT, GkcuFrameParams, host

EXACT_IMAGES = ("OBJEDCTID_0", "EXT_PRIMITIVEID", "ALBEDO", "NORMAL", "MOTIONVECTOR", "MOTIONMOMENT", "PREV_DEPTHBUFFER")
RADIANCE_IMAGES = ("SINGLE_DIFFUSE", "SINGLE_SPECULAR", "ACCUMLATE_DIFFUSE", "ACCUMLATE_SPECULAR", "ACCUMLATE_ALBEDO")


def half_ulp_diff(a, b):
    """|a - b| in units of fp16 ulps (monotone integer mapping of the half bit patterns)."""
    ai = a.view(np.uint16).astype(np.int32)
    bi = b.view(np.uint16).astype(np.int32)
    ai = np.where(ai & 0x8000, 0x8000 - ai, ai)
    bi = np.where(bi & 0x8000, 0x8000 - bi, bi)
    return np.abs(ai - bi)


def mismatch_fraction(a, b):
    ne = a != b
    if ne.ndim == 3:
        ne = ne.any(axis=2)
    return float(ne.mean())


def rel_rmse(a, b):
    a = a.astype(np.float64)
    b = b.astype(np.float64)
    return float(np.sqrt(((a - b) ** 2).mean()) / max(1e-12, np.sqrt((b ** 2).mean())))


def setup_pair(renderer, scene, width, height, sky=None):
    renderer.CreateSwapChain((width, height))
    renderer.OnPostLoadScene(scene, sky)
    osc = O.OracleScene(scene, sky)
    ofr = O.OracleFrame(width, height)
    return osc, ofr


def check_frame(renderer, ofr, exact_fraction=0.0, radiance_fraction=0.0, denoise_ulps=2, images=None):
    """Compares the storage images of the last frame. Integer / G-buffer images must match bit-exactly on all but
    `exact_fraction` of the pixels; radiance images likewise with `radiance_fraction`; RT_DENOISED within `denoise_ulps`
    fp16 ulps (libm pow/exp differ between CUDA and glibc)."""
    report = {}
    for name in images or (EXACT_IMAGES + RADIANCE_IMAGES):
        g, c = renderer.readback(RT[name]), ofr.image(RT[name])
        frac = mismatch_fraction(g, c)
        report[name] = frac
        limit = exact_fraction if name in EXACT_IMAGES else radiance_fraction
        assert frac <= limit, f"{name}: {frac:.3e} of the pixels differ from the oracle (limit {limit:.1e})"
    g, c = renderer.readback(RT["DENOISED"]), ofr.image(RT["DENOISED"])
    ulps = half_ulp_diff(g, c)
    report["DENOISED_max_ulps"] = int(ulps.max())
    bad = float((ulps > denoise_ulps).any(axis=2).mean())
    assert bad <= max(radiance_fraction, exact_fraction), f"RT_DENOISED: {bad:.3e} of the pixels differ by more than {denoise_ulps} fp16 ulps"
    return report


def random_rays(n, seed=1, center=(0, 3, 0), spread=3.0, tmin=1e-3, tmax=1000.0):
    from gknextengine_b200._abi import RAY_DTYPE
    rng = np.random.default_rng(seed)
    rays = np.zeros(n, dtype=RAY_DTYPE)
    o = rng.uniform(-spread, spread, (n, 3)).astype(np.float32) + np.array(center, np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays["ox"], rays["oy"], rays["oz"] = o.T
    rays["dx"], rays["dy"], rays["dz"] = d.T
    rays["tmin"] = tmin
    rays["tmax"] = tmax
    return rays
