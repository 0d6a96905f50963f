"""Python view of the C++ host-side scene code (gknextengine_b200/host/, libgknext_host.so).

Mirrors what the reference's Assets::Scene / NextEngine hand to a renderer each frame
(src/Assets/Scene.cpp:102-289, 443-492; src/Runtime/Engine.cpp:736-849).
"""
import ctypes as C
import os

import numpy as np

from ._abi import GkcuSceneDesc, GkLightObject, GkMaterial, GkNodeProxy, GkSphericalHarmonics, GkUniformBufferObject

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "libgknext_host.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        L = C.CDLL(path)
        L.gkhost_scene_create.restype = C.c_void_p
        L.gkhost_scene_create.argtypes = [C.c_char_p]
        L.gkhost_scene_destroy.argtypes = [C.c_void_p]
        L.gkhost_scene_desc.argtypes = [C.c_void_p, C.POINTER(GkcuSceneDesc)]
        L.gkhost_scene_nodes.restype = C.POINTER(GkNodeProxy)
        L.gkhost_scene_nodes.argtypes = [C.c_void_p, C.POINTER(C.c_uint32)]
        for fn in ("gkhost_scene_provoke_indices", "gkhost_scene_reorders"):
            getattr(L, fn).restype = C.POINTER(C.c_uint32)
            getattr(L, fn).argtypes = [C.c_void_p, C.POINTER(C.c_uint32)]
        L.gkhost_scene_model_indices.restype = C.POINTER(C.c_uint32)
        L.gkhost_scene_model_indices.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32)]
        L.gkhost_scene_materials.restype = C.POINTER(GkMaterial)
        L.gkhost_scene_materials.argtypes = [C.c_void_p, C.POINTER(C.c_uint32)]
        L.gkhost_scene_lights.restype = C.POINTER(GkLightObject)
        L.gkhost_scene_lights.argtypes = [C.c_void_p, C.POINTER(C.c_uint32)]
        L.gkhost_project_sh.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(GkSphericalHarmonics)]
        L.gkhost_scene_changed.restype = C.POINTER(C.c_uint32)
        L.gkhost_scene_changed.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_int)]
        L.gkhost_scene_update_nodes.argtypes = [C.c_void_p]
        L.gkhost_scene_settings.argtypes = [C.c_void_p] + [C.c_int] * 8
        L.gkhost_scene_env.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float]
        L.gkhost_scene_camera.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float]
        L.gkhost_scene_ubo.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_float, C.POINTER(GkUniformBufferObject)]
        L.gkhost_scene_reset_history.argtypes = [C.c_void_p]
        L.gkhost_scene_counts.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)]
        L.gkhost_scene_jitter_instances.argtypes = [C.c_void_p, C.c_uint32, C.c_float]
        L.gkhost_scene_perturb_model.restype = C.c_uint32
        L.gkhost_scene_perturb_model.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
        L.gkhost_make_sky.argtypes = [C.c_uint32, C.c_uint32, C.c_void_p]
        L.gkhost_last_error.restype = C.c_char_p
        L.gkhost_load_hdr.argtypes = [C.c_char_p, C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.gkhost_to_float16.restype = C.c_uint16
        L.gkhost_to_float16.argtypes = [C.c_float]
        L.gkhost_mat4_inverse.argtypes = [C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


class Scene:
    """SceneList::LoadScene + Scene::RebuildMeshBuffer: 'CornellBox.proc', 'RTIO.proc', 'LegoBricks.synth[:N]', or a .glb file (a path, or a
    name looked up in $GKCU_ASSET_DIR and the reference's assets/models; FSceneLoader::LoadGLTFScene's mapping rules, host/gk_gltf.cpp)."""

    def __init__(self, name, samples=8, bounces=5, max_bounces=10, taa=True, temporal_frames=16, denoiser=False, hdr=False,
                 fast_gather=False):
        self._h = lib().gkhost_scene_create(name.encode())
        if not self._h:
            raise ValueError(f"cannot load scene {name!r}: {lib().gkhost_last_error().decode(errors='replace')}")
        self.name = name
        self.settings(samples, bounces, max_bounces, taa, temporal_frames, denoiser, hdr, fast_gather)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().gkhost_scene_destroy(self._h)
            self._h = None

    def settings(self, samples, bounces, max_bounces=10, taa=True, temporal_frames=16, denoiser=False, hdr=False, fast_gather=False):
        self.temporal_frames = temporal_frames
        lib().gkhost_scene_settings(self._h, samples, bounces, max_bounces, int(taa), temporal_frames, int(denoiser), int(hdr), int(fast_gather))

    def env(self, has_sky, has_sun, sky_intensity=100.0, sun_intensity=500.0, sun_rotation=0.5, sky_rotation=0.0):
        lib().gkhost_scene_env(self._h, int(has_sky), int(has_sun), sky_intensity, sun_intensity, sun_rotation, sky_rotation)

    def camera(self, aperture=0.0, focal_distance=10.0, fov=0.0):
        lib().gkhost_scene_camera(self._h, aperture, focal_distance, fov)

    def desc(self):
        d = GkcuSceneDesc()
        lib().gkhost_scene_desc(self._h, C.byref(d))
        return d

    def nodes(self):
        n = C.c_uint32()
        p = lib().gkhost_scene_nodes(self._h, C.byref(n))
        return p, n.value

    def raster_buffers(self):
        """(provoking index buffer, GPUScene.Reorders) of Scene::RebuildMeshBuffer as numpy arrays (Scene.cpp:224-250)."""
        n = C.c_uint32()
        p = lib().gkhost_scene_provoke_indices(self._h, C.byref(n))
        provoke = np.ctypeslib.as_array(p, shape=(n.value,)).copy() if n.value else np.zeros(0, np.uint32)
        p = lib().gkhost_scene_reorders(self._h, C.byref(n))
        reorders = np.ctypeslib.as_array(p, shape=(n.value,)).copy() if n.value else np.zeros(0, np.uint32)
        return provoke, reorders

    def model_indices(self, model):
        """The authored (pre-rotation) model-local index buffer of one model."""
        n = C.c_uint32()
        p = lib().gkhost_scene_model_indices(self._h, model, C.byref(n))
        return np.ctypeslib.as_array(p, shape=(n.value,)).copy() if n.value else np.zeros(0, np.uint32)

    def materials(self):
        n = C.c_uint32()
        p = lib().gkhost_scene_materials(self._h, C.byref(n))
        return p, n.value

    def lights(self):
        n = C.c_uint32()
        p = lib().gkhost_scene_lights(self._h, C.byref(n))
        return p, n.value

    def changed(self):
        """(indices pointer, count, all_changed) of the proxies the last update_nodes() modified."""
        n, a = C.c_uint32(), C.c_int()
        p = lib().gkhost_scene_changed(self._h, C.byref(n), C.byref(a))
        return p, n.value, bool(a.value)

    def update_nodes(self):
        """Scene::UpdateNodes: recompute NodeProxy[] (velocity tick included)."""
        return bool(lib().gkhost_scene_update_nodes(self._h))

    def ubo(self, width, height, total_frames, progressive=False, model_rotation_deg=0.0):
        u = GkUniformBufferObject()
        lib().gkhost_scene_ubo(self._h, width, height, total_frames, int(progressive), model_rotation_deg, C.byref(u))
        return u

    def reset_history(self):
        lib().gkhost_scene_reset_history(self._h)

    def counts(self):
        a, b, c, t = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_uint64()
        lib().gkhost_scene_counts(self._h, C.byref(a), C.byref(b), C.byref(c), C.byref(t))
        return dict(nodes=a.value, models=b.value, materials=c.value, instanced_triangles=t.value)

    def set_node_visible(self, node, visible):
        """Node::SetVisible; the proxy stays in place with visible = 0 (the TLAS instance is masked out)."""
        lib().gkhost_scene_set_node_visible(self._h, node, int(visible))

    def jitter_instances(self, step, fraction):
        lib().gkhost_scene_jitter_instances(self._h, step, fraction)

    def perturb_model(self, model, step):
        n = lib().gkhost_scene_perturb_model(self._h, model, step, None)
        out = np.zeros(n * 4, dtype=np.uint16)
        lib().gkhost_scene_perturb_model(self._h, model, step, out.ctypes.data)
        return out


def make_sky(width=1024, height=512):
    """Procedural stand-in for assets/textures/river_road_2.hdr (sky id 0, Engine.cpp:855): RGBA32F equirect."""
    out = np.zeros((height, width, 4), dtype=np.float32)
    lib().gkhost_make_sky(width, height, out.ctypes.data)
    return out


def load_hdr(path):
    """A Radiance .hdr equirect picture as (H, W, 4) float32, rows top to bottom - what stbi_loadf hands to the reference's texture pool
    (src/Assets/Texture.cpp:390-432) and what gkcu_set_sky takes."""
    w, h = C.c_uint32(), C.c_uint32()
    if lib().gkhost_load_hdr(path.encode(), None, C.byref(w), C.byref(h)):
        raise ValueError(f"{path}: {lib().gkhost_last_error().decode(errors='replace')}")
    out = np.zeros((h.value, w.value, 4), dtype=np.float32)
    lib().gkhost_load_hdr(path.encode(), out.ctypes.data, C.byref(w), C.byref(h))
    return out


def project_sh(rgba):
    """GlobalTexturePool's ProjectHdrToSh (src/Assets/Texture.cpp:180-262) of an equirect RGBA32F image -> GkSphericalHarmonics."""
    rgba = np.ascontiguousarray(rgba, dtype=np.float32)
    out = GkSphericalHarmonics()
    lib().gkhost_project_sh(rgba.ctypes.data, rgba.shape[1], rgba.shape[0], C.byref(out))
    return out
