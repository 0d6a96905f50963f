"""ctypes wrapper of oracle/liboracle.so (and oracle/_ref/liboracle_tinybvh.so). TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never by the product."""
import ctypes as C
import os

import numpy as np

from gknextengine_b200._abi import (RT_FORMAT, GkcuCounters, GkcuFrameParams, GkcuSceneDesc, GkMaterial, GkNodeProxy,
                                    GkUniformBufferObject, HIT_DTYPE)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NATIVE = os.path.join(ROOT, "oracle", "liboracle.so")
TINYBVH = os.path.join(ROOT, "oracle", "_ref", "liboracle_tinybvh.so")
_libs = {}


def load(path=NATIVE):
    if path in _libs:
        return _libs[path]
    if not os.path.exists(path):
        raise RuntimeError(f"{path} missing: run `make -C oracle` (or __graft_entry__.build())")
    L = C.CDLL(path)
    L.gko_scene_create.restype = C.c_void_p
    L.gko_scene_create.argtypes = [C.POINTER(GkcuSceneDesc)]
    L.gko_scene_destroy.argtypes = [C.c_void_p]
    L.gko_update_nodes.argtypes = [C.c_void_p, C.POINTER(GkNodeProxy), C.c_uint32]
    L.gko_update_materials.argtypes = [C.c_void_p, C.POINTER(GkMaterial), C.c_uint32]
    L.gko_refit_model.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32]
    L.gko_set_sky.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32]
    L.gko_set_probes.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.gko_set_lights.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
    L.gko_set_sky_sh.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
    L.gko_probes_voxelize.argtypes = [C.c_void_p, C.c_int]
    L.gko_probes_bake.argtypes = [C.c_void_p, C.POINTER(GkUniformBufferObject), C.c_uint32, C.c_uint32, C.c_int]
    for fn in ('gko_probes_cubes', 'gko_probes_voxels', 'gko_probes_pages'):
        getattr(L, fn).restype = C.c_void_p
        getattr(L, fn).argtypes = [C.c_void_p]
    L.gko_pack_rgb10a2.restype = C.c_uint32
    L.gko_pack_rgb10a2.argtypes = [C.POINTER(C.c_float * 4)]
    L.gko_unpack_rgb10a2.argtypes = [C.c_uint32, C.POINTER(C.c_float * 4)]
    L.gko_sample_ibl_rough.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_float * 3), C.c_float, C.POINTER(C.c_float * 4)]
    L.gko_set_raster_buffers.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32]
    L.gko_cull.argtypes = [C.c_void_p, C.POINTER(GkUniformBufferObject), C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
    L.gko_visibility_pass.argtypes = [C.c_void_p, C.POINTER(GkUniformBufferObject), C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]
    L.gko_intersector_name.restype = C.c_char_p
    L.gko_frame_create.restype = C.c_void_p
    L.gko_frame_create.argtypes = [C.c_uint32, C.c_uint32]
    L.gko_frame_destroy.argtypes = [C.c_void_p]
    L.gko_render_frame.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(GkUniformBufferObject), C.POINTER(GkcuFrameParams), C.c_int,
                                   C.c_uint32, C.c_uint32, C.c_int]
    L.gko_reproject.argtypes = [C.c_void_p, C.POINTER(GkcuFrameParams)]
    L.gko_denoise.argtypes = [C.c_void_p, C.POINTER(GkUniformBufferObject)]
    L.gko_end_frame_copies.argtypes = [C.c_void_p]
    L.gko_image.restype = C.c_void_p
    L.gko_image.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_size_t)]
    L.gko_get_counters.argtypes = [C.c_void_p, C.POINTER(GkcuCounters)]
    L.gko_last_rt_seconds.restype = C.c_double
    L.gko_last_rt_seconds.argtypes = [C.c_void_p]
    L.gko_trace_rays.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32]
    L.gko_pcg4d.argtypes = [C.POINTER(C.c_uint32 * 4)]
    L.gko_uint_to_float.restype = C.c_float
    L.gko_uint_to_float.argtypes = [C.c_uint32]
    L.gko_float_to_half.restype = C.c_uint16
    L.gko_float_to_half.argtypes = [C.c_float]
    L.gko_half_to_float.restype = C.c_float
    L.gko_half_to_float.argtypes = [C.c_uint16]
    L.gko_gt_tonemap.argtypes = [C.POINTER(C.c_float * 3), C.POINTER(C.c_float * 3)]
    L.gko_ggx_sampling.argtypes = [C.POINTER(C.c_uint32 * 4), C.c_float, C.POINTER(C.c_float * 3), C.POINTER(C.c_float * 3)]
    L.gko_random_in_hemisphere1.argtypes = [C.POINTER(C.c_uint32 * 4), C.POINTER(C.c_float * 3)]
    L.gko_det_sincos.argtypes = [C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.gko_radiance.restype = C.POINTER(C.c_float)
    L.gko_radiance.argtypes = [C.c_void_p, C.POINTER(C.c_size_t)]
    L.gko_resolve.argtypes = [C.c_void_p, C.POINTER(GkUniformBufferObject)]
    L.gko_schlick.restype = C.c_float
    L.gko_schlick.argtypes = [C.c_float, C.c_float]
    _libs[path] = L
    return L


class OracleScene:
    def __init__(self, scene, sky=None, lib_path=NATIVE):
        self.L = load(lib_path)
        d = scene.desc()
        self._h = self.L.gko_scene_create(C.byref(d))
        mats, nm = scene.materials()
        self.L.gko_update_materials(self._h, mats, nm)
        self.update_nodes(scene)
        if hasattr(scene, "raster_buffers"):
            provoke, reorders = scene.raster_buffers()
            self.L.gko_set_raster_buffers(self._h, provoke.ctypes.data, len(provoke), reorders.ctypes.data, len(reorders))
        if sky is not None:
            self.set_sky(0, sky)

    def __del__(self):
        if getattr(self, "_h", None):
            self.L.gko_scene_destroy(self._h)
            self._h = None

    def update_nodes(self, scene):
        nodes, nn = scene.nodes()
        self.L.gko_update_nodes(self._h, nodes, nn)
        self._node_count = nn

    # ---- visibility front end
    def cull(self, ubo, prev_depth):
        """(draw commands, GPUDrivenStat dict) of Task.GpuCull against the depth image `prev_depth` (H, W) float32."""
        from gknextengine_b200._abi import DRAW_CMD_DTYPE
        prev_depth = np.ascontiguousarray(prev_depth, dtype=np.float32)
        cmds = np.zeros(max(1, self._node_count), DRAW_CMD_DTYPE)
        stats = np.zeros(8, np.uint32)
        self.L.gko_cull(self._h, C.byref(ubo), prev_depth.ctypes.data, prev_depth.shape[1], prev_depth.shape[0], cmds.ctypes.data, stats.ctypes.data)
        names = ("ProcessedCount", "VisibleCount", "CulledCount", "TriangleCount", "CulledTriangleCount")
        return cmds[:self._node_count], {n: int(stats[i]) for i, n in enumerate(names)}

    def visibility_pass(self, ubo, cmds, width, height):
        """RT_MINIGBUFFER (H, W) uint32 of the software raster of the draw commands."""
        cmds = np.ascontiguousarray(cmds)
        vis = np.zeros((height, width), np.uint32)
        self.L.gko_visibility_pass(self._h, C.byref(ubo), cmds.ctypes.data, len(cmds), width, height, vis.ctypes.data)
        return vis

    def set_sky(self, idx, rgba):
        rgba = np.ascontiguousarray(rgba, dtype=np.float32)
        self.L.gko_set_sky(self._h, idx, rgba.ctypes.data, rgba.shape[1], rgba.shape[0])

    # ---- probe grid
    def set_lights(self, scene):
        p, n = scene.lights()
        self.L.gko_set_lights(self._h, C.cast(p, C.c_void_p), n)

    def set_sky_sh(self, idx, sh):
        self.L.gko_set_sky_sh(self._h, idx, C.byref(sh))

    def set_probes(self, cubes, voxels):
        self.L.gko_set_probes(self._h, cubes.ctypes.data, voxels.ctypes.data)

    def probes_voxelize(self, threads=0):
        self.L.gko_probes_voxelize(self._h, threads)

    def probes_bake(self, ubo, first, count, threads=0):
        self.L.gko_probes_bake(self._h, C.byref(ubo), first, count, threads)

    def probes_read(self, pages=False):
        from gknextengine_b200._abi import GK_CUBE_COUNT, GK_PAGE_COUNT
        from gknextengine_b200.renderer import CudaPathTracingRenderer as R
        def view(ptr, n, dt):
            return np.frombuffer((C.c_uint8 * (n * dt.itemsize)).from_address(ptr), dtype=dt).copy()
        c = view(self.L.gko_probes_cubes(self._h), GK_CUBE_COUNT, R.CUBE_DTYPE)
        v = view(self.L.gko_probes_voxels(self._h), GK_CUBE_COUNT, R.VOXEL_DTYPE)
        pg = view(self.L.gko_probes_pages(self._h), GK_PAGE_COUNT * GK_PAGE_COUNT, R.PAGE_DTYPE) if pages else None
        return c, v, pg

    def refit_model(self, model, simple_vertices):
        self.L.gko_refit_model(self._h, model, simple_vertices.ctypes.data, len(simple_vertices) // 4)

    def trace(self, rays, any_hit=False):
        rays = np.ascontiguousarray(rays)
        hits = np.zeros(len(rays), dtype=HIT_DTYPE)
        self.L.gko_trace_rays(self._h, rays.ctypes.data, len(rays), hits.ctypes.data, int(any_hit))
        return hits


class OracleFrame:
    def __init__(self, width, height, lib_path=NATIVE):
        self.L = load(lib_path)
        self.w, self.h = width, height
        self._h = self.L.gko_frame_create(width, height)

    def __del__(self):
        if getattr(self, "_h", None):
            self.L.gko_frame_destroy(self._h)
            self._h = None

    def render(self, oscene, ubo, params, threads=0, rows=None, run_post=True):
        y0, y1 = rows if rows else (0, self.h)
        self.L.gko_render_frame(oscene._h, self._h, C.byref(ubo), C.byref(params), threads, y0, y1, int(run_post))

    def reproject(self, params):
        self.L.gko_reproject(self._h, C.byref(params))

    def denoise(self, ubo):
        self.L.gko_denoise(self._h, C.byref(ubo))

    def end_frame_copies(self):
        self.L.gko_end_frame_copies(self._h)

    def image(self, rt):
        """Writable numpy view of a storage image (H, W, C) in its reference format."""
        n = C.c_size_t()
        p = self.L.gko_image(self._h, rt, C.byref(n))
        dtype, ch = RT_FORMAT[rt]
        buf = (C.c_char * n.value).from_address(p)
        a = np.frombuffer(buf, dtype=dtype)
        return a.reshape(self.h, self.w, ch) if ch > 1 else a.reshape(self.h, self.w)

    def radiance(self):
        """Writable view of the multi-GPU radiance planes: (3, H, W, 4) float32 = FinalColor sums, FinalReflection sums, direct."""
        n = C.c_size_t()
        p = self.L.gko_radiance(self._h, C.byref(n))
        return np.ctypeslib.as_array(p, shape=(n.value,)).reshape(3, self.h, self.w, 4)

    def resolve(self, ubo):
        self.L.gko_resolve(self._h, C.byref(ubo))

    def counters(self):
        c = GkcuCounters()
        self.L.gko_get_counters(self._h, C.byref(c))
        return dict(primary=c.primary_rays, closest=c.closest_hit_rays, any=c.any_hit_rays)

    def rt_seconds(self):
        return self.L.gko_last_rt_seconds(self._h)
