"""Shared helpers of the parity tests: build a scene, render it with the CUDA renderer and with the CPU oracle, compare."""
import os

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


def denoised_ulps(g, c):
    """fp16 ulp distance per channel of two RT_DENOISED images; NaN on both sides (0 / 0 of the bilateral weights next to a light, a zero
    refraction direction) counts as equal - the two processors spell NaN differently."""
    ulps = half_ulp_diff(g, c)
    ulps[np.isnan(g.astype(np.float32)) & np.isnan(c.astype(np.float32))] = 0
    return ulps


def mismatch_fraction(a, b):
    """Fraction of pixels that differ; a NaN on both sides counts as equal (the two processors spell NaN differently)."""
    ne = a != b
    if a.dtype.kind == "f":
        ne &= ~(np.isnan(a) & np.isnan(b))
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


def check_frame(renderer, ofr, exact_fraction=0.0, radiance_fraction=0.0, denoise_ulps=0, images=None):
    """Compares the storage images of the last frame with the oracle's. Since the leaf transcendentals are pinned too (include/gk_detmath.h)
    every image, RT_DENOISED included, is expected bit-identical; the fraction arguments remain for experiments and default to 0."""
    report = {}
    for name in images or (EXACT_IMAGES + RADIANCE_IMAGES):
        g, c = renderer.readback(RT[name]), ofr.image(RT[name])
        frac = mismatch_fraction(g, c)
        report[name] = frac
        limit = exact_fraction if name in EXACT_IMAGES else radiance_fraction
        assert frac <= limit, f"{name}: {frac:.3e} of the pixels differ from the oracle (limit {limit:.1e})"
    ulps = denoised_ulps(renderer.readback(RT["DENOISED"]), ofr.image(RT["DENOISED"]))
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


class CompiledReferenceFrame:
    """The reference's own frame - Util.BufferClear, Core.PathTracing, Process.ReProject, Process.DenoiseJBF and the copy pass - run from its
    shader text compiled as C++ (oracle/_ref/libref_pathtrace.so + libref_stencils.so, built by `make -C oracle ref` when the reference is
    mounted). The ray query, the texture sampler and the visibility buffer are served by the oracle scene `osc` (hardware in the reference)."""
    NAMES16 = ("SINGLE_DIFFUSE", "SINGLE_SPECULAR", "ALBEDO", "NORMAL", "DENOISED", "ACCUMLATE_DIFFUSE", "ACCUMLATE_SPECULAR", "ACCUMLATE_ALBEDO",
               "SINGLE_PREV_DIFFUSE", "SINGLE_PREV_SPECULAR", "SINGLE_PREV_ALBEDO")

    @staticmethod
    def available():
        import os
        d = os.path.dirname(O.TINYBVH)
        return all(os.path.exists(os.path.join(d, n)) for n in ("libref_pathtrace.so", "libref_stencils.so"))

    def __init__(self, width, height):
        import ctypes as C
        import os
        d = os.path.dirname(O.TINYBVH)
        self.pt, self.st = C.CDLL(os.path.join(d, "libref_pathtrace.so")), C.CDLL(os.path.join(d, "libref_stencils.so"))
        self.pt.ref_pathtrace.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_int, C.POINTER(C.c_uint64)]
        self.st.ref_reproject.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        self.st.ref_denoise.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_void_p]
        self.w, self.h = width, height
        W, H = width, height
        self.images = {n: np.zeros((H, W, 4), np.float16) for n in self.NAMES16}
        self.images.update({"OBJEDCTID_0": np.zeros((H, W), np.uint32), "OBJEDCTID_1": np.zeros((H, W), np.uint32), "MINIGBUFFER": np.zeros((H, W), np.uint32),
                            "MOTIONVECTOR": np.zeros((H, W, 2), np.float32), "PREV_DEPTHBUFFER": np.zeros((H, W), np.float32),
                            "MOTIONMOMENT": np.zeros((H, W), np.uint16)})
        self.table = (C.c_void_p * 64)()
        for k, v in self.images.items():
            self.table[RT[k]] = v.ctypes.data
        self.closest = self.any = 0

    def render(self, osc, ubo, temporal_frames=16, progressive=0, fast_reproject=0):
        import ctypes as C
        osc.L.gko_cast_visibility.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
        osc.L.gko_cast_visibility(osc._h, C.byref(ubo), self.w, self.h, self.images["MINIGBUFFER"].ctypes.data)  # stands in for the raster pass
        counters = (C.c_uint64 * 2)()
        self.pt.ref_pathtrace(osc._h, C.byref(ubo), self.table, self.w, self.h, counters)
        self.closest, self.any = counters[0], counters[1]
        self.st.ref_reproject(self.table, self.w, self.h, progressive, temporal_frames, fast_reproject)
        self.st.ref_denoise(self.table, self.w, self.h, C.cast(C.byref(ubo), C.c_void_p))

    def end_frame(self):
        """The copy pass and the object-id copy (PathTracingRenderer.cpp:160-182, VulkanBaseRenderer.cpp:1136-1152)."""
        for acc, prev in (("ACCUMLATE_DIFFUSE", "SINGLE_PREV_DIFFUSE"), ("ACCUMLATE_SPECULAR", "SINGLE_PREV_SPECULAR"), ("ACCUMLATE_ALBEDO", "SINGLE_PREV_ALBEDO")):
            self.images[prev][...] = self.images[acc]
        self.images["OBJEDCTID_1"][...] = self.images["OBJEDCTID_0"]

    def mismatches(self, name, other):
        """Pixels whose bit patterns differ between this frame's image `name` and `other`."""
        a = self.images[name]
        ua, ub = (x.view(np.uint16 if x.dtype.itemsize == 2 else np.uint32).reshape(self.h, self.w, -1) for x in (a, other))
        return (ua != ub).any(axis=2)


def textured_cornell_materials(scene, rng):
    """Three random textures (albedo, MRA, tangent-space normals) bound to every non-emissive Cornell material; returns ((id, image), ...)."""
    nrm = rng.normal(size=(32, 32, 3)) * 0.35 + np.array([0, 0, 1.0])
    nrm /= np.linalg.norm(nrm, axis=2, keepdims=True)
    textures = ((1, rng.uniform(0.2, 1.0, (16, 16, 4)).astype(np.float32)), (2, rng.uniform(0.0, 1.0, (8, 32, 4)).astype(np.float32)),
                (3, np.concatenate([nrm * 0.5 + 0.5, np.ones((32, 32, 1))], axis=2).astype(np.float32)))
    mats, nm = scene.materials()
    for i in range(nm):
        if mats[i].MaterialModel != 4:
            mats[i].DiffuseTextureId, mats[i].MRATextureId, mats[i].NormalTextureId = 1, 2, 3
            mats[i].Fuzziness = 0.8
            mats[i].Metalness = 0.5 if i % 2 else 0.0
    return textures, mats, nm
