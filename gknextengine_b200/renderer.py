"""Python mirror of the reference's renderer plug-in interface for the CUDA path tracer.

Reference interface (src/Rendering/VulkanBaseRenderer.hpp:45-51, 213-243; .cpp:892-917):
    enum ERendererType { ERT_PathTracing, ERT_ModernDeferred, ERT_LegacyDeferred, ERT_VoxelTracing }
    class LogicRendererBase { OnDeviceSet(); CreateSwapChain(extent); DeleteSwapChain(); Render(cmd, imageIndex); BeforeNextFrame(); }
    VulkanBaseRenderer::RegisterLogicRenderer(type) / SwitchLogicRenderer(type)
and the acceleration-structure hooks of RayTraceBaseRenderer (src/Rendering/RayTraceBaseRenderer.cpp:165-218):
    OnPreLoadScene / OnPostLoadScene / AfterUpdateScene / PreRender.

Everything below calls the C ABI of libgknext_cuda.so (include/gkcu.h) through ctypes; there is no CPU fallback:
constructing a renderer without the library or without a B200 raises.
"""
import ctypes as C
import enum
import os

import numpy as np

from ._abi import (GK_CUBE_COUNT, GK_PAGE_COUNT, GkLightObject, GkSphericalHarmonics, HIT_DTYPE, RAY_DTYPE, RT, RT_FORMAT, GkcuCounters, GkcuKernelStats, GkcuFrameParams, GkcuSceneDesc, GkcuTimings, GkMaterial,
                   GkNodeProxy, GkUniformBufferObject)

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.environ.get("GKCU_LIB") or os.path.join(_HERE, "libgknext_cuda.so")  # GKCU_LIB: A/B runs of two builds
_lib = None


class GkcuError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise GkcuError(f"{_LIB_PATH} is missing; build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(there is no CPU fallback)")
        L = C.CDLL(_LIB_PATH)
        P = C.c_void_p
        L.gkcu_last_error.restype = C.c_char_p
        L.gkcu_create.argtypes = [C.c_int, C.POINTER(P)]
        L.gkcu_destroy.argtypes = [P]
        L.gkcu_resize.argtypes = [P, C.c_uint32, C.c_uint32]
        L.gkcu_scene_upload_raster.argtypes = [P, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32]
        L.gkcu_cull.argtypes = [P, C.POINTER(GkUniformBufferObject)]
        L.gkcu_visibility_pass.argtypes = [P, C.POINTER(GkUniformBufferObject)]
        L.gkcu_get_gpu_driven_stats.argtypes = [P, C.c_void_p]
        L.gkcu_read_draw_commands.argtypes = [P, C.c_void_p, C.c_uint32]
        L.gkcu_bind_external_image.argtypes = [P, C.c_uint32, C.c_uint32, C.c_int, C.c_size_t, C.c_size_t, C.c_uint32]
        L.gkcu_unbind_external_image.argtypes = [P, C.c_uint32]
        L.gkcu_bind_external_semaphore.argtypes = [P, C.c_uint32, C.c_int]
        L.gkcu_scene_upload.argtypes = [P, C.POINTER(GkcuSceneDesc)]
        L.gkcu_blas_build.argtypes = [P]
        L.gkcu_blas_refit.argtypes = [P, C.c_uint32, C.c_void_p, C.c_uint32]
        L.gkcu_update_nodes.argtypes = [P, C.POINTER(GkNodeProxy), C.c_uint32]
        L.gkcu_update_nodes_sparse.argtypes = [P, C.POINTER(C.c_uint32), C.c_void_p, C.c_uint32]
        L.gkcu_update_materials.argtypes = [P, C.POINTER(GkMaterial), C.c_uint32]
        L.gkcu_tlas_build.argtypes = [P]
        L.gkcu_set_sky.argtypes = [P, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32]
        L.gkcu_set_probes.argtypes = [P, C.c_void_p, C.c_void_p]
        L.gkcu_set_lights.argtypes = [P, C.c_void_p, C.c_uint32]
        L.gkcu_set_sky_sh.argtypes = [P, C.c_uint32, C.POINTER(GkSphericalHarmonics)]
        L.gkcu_probes_voxelize.argtypes = [P]
        L.gkcu_probes_bake.argtypes = [P, C.POINTER(GkUniformBufferObject), C.c_uint32, C.c_uint32]
        L.gkcu_probes_read.argtypes = [P, C.c_void_p, C.c_void_p, C.c_void_p]
        L.gkcu_peer_export.argtypes = [P, C.c_uint32, C.c_void_p, C.POINTER(C.c_void_p)]
        L.gkcu_peer_import.argtypes = [P, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
        L.gkcu_render_resolve_post_strip.argtypes = [P, C.POINTER(GkUniformBufferObject), C.POINTER(GkcuFrameParams)]
        L.gkcu_render_resolve_strip.argtypes = [P, C.POINTER(GkUniformBufferObject), C.POINTER(GkcuFrameParams)]
        L.gkcu_render_frame.argtypes = [P, C.POINTER(GkUniformBufferObject), C.POINTER(GkcuFrameParams)]
        L.gkcu_radiance_buffer.argtypes = [P, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
        L.gkcu_render_resolve.argtypes = [P, C.POINTER(GkUniformBufferObject), C.POINTER(GkcuFrameParams)]
        L.gkcu_readback.argtypes = [P, C.c_uint32, C.c_void_p, C.c_size_t]
        L.gkcu_readback_denoised_async.argtypes = [P, C.c_void_p, C.c_size_t]
        L.gkcu_readback_wait.argtypes = [P]
        L.gkcu_upload_image.argtypes = [P, C.c_uint32, C.c_void_p, C.c_size_t]
        L.gkcu_run_post.argtypes = [P, C.POINTER(GkUniformBufferObject), C.POINTER(GkcuFrameParams)]
        L.gkcu_debug_read.argtypes = [P, C.c_uint32, C.c_size_t, C.c_void_p, C.c_size_t]
        L.gkcu_debug_info.argtypes = [P, C.POINTER(C.c_uint32), C.c_uint32]
        L.gkcu_synchronize.argtypes = [P]
        L.gkcu_get_timings.argtypes = [P, C.POINTER(GkcuTimings)]
        L.gkcu_get_counters.argtypes = [P, C.POINTER(GkcuCounters)]
        L.gkcu_get_kernel_stats.argtypes = [P, C.POINTER(GkcuKernelStats)]
        L.gkcu_set_kernel_timing.argtypes = [P, C.c_uint32]
        L.gkcu_set_fast_stencil.argtypes = [P, C.c_uint32]
        L.gkcu_set_skip_unchanged_nodes.argtypes = [P, C.c_uint32]
        L.gkcu_set_tlas_refit.argtypes = [P, C.c_uint32, C.c_uint32]
        L.gkcu_tlas_stats.argtypes = [P, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.gkcu_stream.restype = C.c_uint64
        L.gkcu_stream.argtypes = [P]
        L.gkcu_trace_rays.argtypes = [P, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32]
        _lib = L
    return _lib


def _ck(rc):
    if rc != 0:
        raise GkcuError(lib().gkcu_last_error().decode(errors="replace"))


class ERendererType(enum.IntEnum):
    """VulkanBaseRenderer.hpp:45-51 plus the new entry the CUDA renderer registers under."""
    ERT_PathTracing = 0
    ERT_ModernDeferred = 1
    ERT_LegacyDeferred = 2
    ERT_VoxelTracing = 3
    ERT_CudaPathTracing = 4


class CudaPathTracingRenderer:
    """LogicRendererBase implementation backed by libgknext_cuda.so (the drop-in for RayTracing::PathTracingRenderer)."""

    def __init__(self, device=0):
        self.device = device
        self._ctx = C.c_void_p()
        self.extent = None
        self.OnDeviceSet()

    # ---- LogicRendererBase
    def OnDeviceSet(self):
        if not self._ctx:
            _ck(lib().gkcu_create(self.device, C.byref(self._ctx)))

    def CreateSwapChain(self, extent):
        w, h = extent
        _ck(lib().gkcu_resize(self._ctx, w, h))
        self.extent = (w, h)

    def DeleteSwapChain(self):
        self.extent = None

    def Render(self, ubo, params=None):
        """PathTracingRenderer::Render: rt pass -> reproject pass -> compose pass -> copy pass (asynchronous)."""
        params = params or GkcuFrameParams(temporal_frames=16, spp_world=1)
        _ck(lib().gkcu_render_frame(self._ctx, C.byref(ubo), C.byref(params)))

    def BeforeNextFrame(self):
        _ck(lib().gkcu_synchronize(self._ctx))

    # ---- RayTraceBaseRenderer hooks
    def OnPostLoadScene(self, scene, sky=None):
        """Scene::RebuildMeshBuffer uploads + CreateAccelerationStructures."""
        d = scene.desc()
        _ck(lib().gkcu_scene_upload(self._ctx, C.byref(d)))
        mats, nm = scene.materials()
        _ck(lib().gkcu_update_materials(self._ctx, mats, nm))
        _ck(lib().gkcu_blas_build(self._ctx))
        if hasattr(scene, "raster_buffers"):  # the buffers only the visibility pass reads (Scene.cpp:240-250)
            provoke, reorders = scene.raster_buffers()
            _ck(lib().gkcu_scene_upload_raster(self._ctx, provoke.ctypes.data, len(provoke), reorders.ctypes.data, len(reorders)))
        self.AfterUpdateScene(scene)
        if sky is not None:
            self.set_sky(0, sky)

    # ---- visibility front end (VulkanBaseRenderer::PreRender: "gpu cull", "visibility pass")
    def cull(self, ubo):
        _ck(lib().gkcu_cull(self._ctx, C.byref(ubo)))

    def visibility_pass(self, ubo):
        _ck(lib().gkcu_visibility_pass(self._ctx, C.byref(ubo)))

    def gpu_driven_stats(self):
        from ._abi import GkGPUDrivenStat
        st = GkGPUDrivenStat()
        _ck(lib().gkcu_get_gpu_driven_stats(self._ctx, C.byref(st)))
        return {n: getattr(st, n) for n in ("ProcessedCount", "VisibleCount", "CulledCount", "TriangleCount", "CulledTriangleCount")}

    def draw_commands(self, count):
        from ._abi import DRAW_CMD_DTYPE
        out = np.zeros(count, DRAW_CMD_DTYPE)
        _ck(lib().gkcu_read_draw_commands(self._ctx, out.ctypes.data, count))
        return out

    def AfterUpdateScene(self, scene):
        nodes, nn = scene.nodes()
        _ck(lib().gkcu_update_nodes(self._ctx, nodes, nn))

    def AfterUpdateSceneSparse(self, scene):
        """Like AfterUpdateScene, but uploads only the proxies the last Scene.update_nodes() changed (gkcu_update_nodes_sparse)."""
        idx, n, all_changed = scene.changed()
        if all_changed:
            return self.AfterUpdateScene(scene)
        if n == 0:
            return
        nodes, _ = scene.nodes()
        ids = np.ctypeslib.as_array(idx, shape=(n,)).copy()
        all_nodes = np.ctypeslib.as_array(C.cast(nodes, C.POINTER(C.c_uint8)), shape=(int(ids.max()) + 1, C.sizeof(GkNodeProxy)))
        packed = np.ascontiguousarray(all_nodes[ids])
        _ck(lib().gkcu_update_nodes_sparse(self._ctx, ids.ctypes.data_as(C.POINTER(C.c_uint32)), packed.ctypes.data, n))

    def PreRender(self):
        _ck(lib().gkcu_tlas_build(self._ctx))

    def update_materials(self, scene):
        mats, nm = scene.materials()
        _ck(lib().gkcu_update_materials(self._ctx, mats, nm))

    def refit_model(self, model, simple_vertices):
        sv = np.ascontiguousarray(simple_vertices, dtype=np.uint16)
        _ck(lib().gkcu_blas_refit(self._ctx, model, sv.ctypes.data, len(sv) // 4))

    def set_sky(self, idx, rgba):
        rgba = np.ascontiguousarray(rgba, dtype=np.float32)
        _ck(lib().gkcu_set_sky(self._ctx, idx, rgba.ctypes.data, rgba.shape[1], rgba.shape[0]))

    # ---- ambient-cube probe grid
    CUBE_DTYPE = np.dtype([(n, "<u4") for n in ("PosZ", "NegZ", "PosY", "NegY", "PosX", "NegX", "PosZ_D", "NegZ_D", "PosY_D", "NegY_D", "PosX_D", "NegX_D",
                                                "skyVisibility_pznzpyny", "skyVisibility_pxnxs0s1")])
    VOXEL_DTYPE = np.dtype([(n, "<u4") for n in ("matId", "age", "distanceToSolid_gg_z01", "distanceToSolid_x01_y01")])
    PAGE_DTYPE = np.dtype([(n, "<u4") for n in ("pageId", "voxelCount", "voxelDataIdx", "probeDataIdx")])

    def set_lights(self, scene):
        """GPUScene.Lights of the loaded scene."""
        p, n = scene.lights()
        _ck(lib().gkcu_set_lights(self._ctx, C.cast(p, C.c_void_p), n))

    def set_sky_sh(self, idx, sh):
        """GPUScene.HDRSHs[idx] (host.project_sh of the sky image)."""
        _ck(lib().gkcu_set_sky_sh(self._ctx, idx, C.byref(sh)))

    def probes_voxelize(self):
        """FCPUAccelerationStructure::AsyncProcessFull + FCPUPageIndex::UpdateData over the whole grid, on the GPU."""
        _ck(lib().gkcu_probes_voxelize(self._ctx))

    def probes_bake(self, ubo, first_cube, cube_count):
        """One Bake.HwAmbientCube dispatch over probes [first_cube, first_cube + cube_count)."""
        _ck(lib().gkcu_probes_bake(self._ctx, C.byref(ubo), first_cube, cube_count))

    BAKE_TEMPORAL_FRAMES = {0: 30, 1: 120}  # UserSettings.BakeSpeedLevel, RayTraceBaseRenderer.cpp:237-251 (level 2 = off)

    def PostRender(self, ubo, frame_count, bake_speed_level=1):
        """RayTraceBaseRenderer::PostRender (RayTraceBaseRenderer.cpp:225-274): bake 1/temporalFrames of the grid per frame."""
        tf = self.BAKE_TEMPORAL_FRAMES.get(bake_speed_level)
        if tf is None:
            return
        group = GK_CUBE_COUNT // 64
        group_per_frame = group // tf
        self.probes_bake(ubo, (frame_count % tf) * group_per_frame * 64, group_per_frame * 64)

    def probes_read(self, cubes=True, voxels=True, pages=False):
        c = np.zeros(GK_CUBE_COUNT, dtype=self.CUBE_DTYPE) if cubes else None
        v = np.zeros(GK_CUBE_COUNT, dtype=self.VOXEL_DTYPE) if voxels else None
        pg = np.zeros(GK_PAGE_COUNT * GK_PAGE_COUNT, dtype=self.PAGE_DTYPE) if pages else None
        _ck(lib().gkcu_probes_read(self._ctx, c.ctypes.data if cubes else None, v.ctypes.data if voxels else None, pg.ctypes.data if pages else None))
        return c, v, pg

    # ---- multi-GPU split
    def radiance_buffer(self):
        p, n = C.c_void_p(), C.c_size_t()
        _ck(lib().gkcu_radiance_buffer(self._ctx, C.byref(p), C.byref(n)))
        return p.value, n.value

    PEER_RADIANCE, PEER_DIRECT, PEER_SINGLE_DIFFUSE, PEER_SINGLE_SPECULAR, PEER_FLAGS = 0, 1, 2, 3, 4
    PEER_ACCUM = (5, 6, 7, 8, 9, 10)  # both halves of the ACCUM_DIFFUSE / SPECULAR / ALBEDO ping-pong pairs
    PEER_DENOISED = 11
    PEER_KINDS = 12

    def peer_export(self, what):
        """(64-byte cudaIpcMemHandle, raw device pointer) of one of this rank's buffers."""
        h = (C.c_uint8 * 64)()
        p = C.c_void_p()
        _ck(lib().gkcu_peer_export(self._ctx, what, h, C.byref(p)))
        return bytes(h), p.value

    def peer_import(self, rank, what, handle=None, ptr=None):
        """Registers rank `rank`'s buffer: `handle` from another process, or `ptr` from a context of this process."""
        hb = (C.c_uint8 * 64).from_buffer_copy(handle) if handle is not None else None
        _ck(lib().gkcu_peer_import(self._ctx, rank, what, hb, C.c_void_p(ptr) if ptr else None))

    def ResolveStrip(self, ubo, params):
        """Fused reduce + resolve of this rank's strip over peer memory. With every rank's PEER_FLAGS imported the two cross-rank
        barriers run on the device; otherwise the caller brackets this call with its own stream-ordered barriers."""
        _ck(lib().gkcu_render_resolve_strip(self._ctx, C.byref(ubo), C.byref(params)))

    def ResolvePostStrip(self, ubo, params):
        """The whole tail of a multi-GPU frame on every rank: strip resolve + strip reproject (history mirrored to every rank) + strip compose
        into rank 0's RT_DENOISED, between two device barriers (gkcu_render_resolve_post_strip)."""
        _ck(lib().gkcu_render_resolve_post_strip(self._ctx, C.byref(ubo), C.byref(params)))

    def Resolve(self, ubo, params):
        _ck(lib().gkcu_render_resolve(self._ctx, C.byref(ubo), C.byref(params)))

    # ---- results
    def readback(self, rt, out=None):
        w, h = self.extent
        dtype, ch = RT_FORMAT[rt]
        shape = (h, w, ch) if ch > 1 else (h, w)
        if out is None:
            out = np.empty(shape, dtype=dtype)
        _ck(lib().gkcu_readback(self._ctx, rt, out.ctypes.data, out.nbytes))
        return out

    def upload_image(self, rt, arr):
        arr = np.ascontiguousarray(arr)
        _ck(lib().gkcu_upload_image(self._ctx, rt, arr.ctypes.data, arr.nbytes))

    def run_post(self, ubo, params):
        _ck(lib().gkcu_run_post(self._ctx, C.byref(ubo), C.byref(params)))

    def timings(self):
        t = GkcuTimings()
        _ck(lib().gkcu_get_timings(self._ctx, C.byref(t)))
        return {n: getattr(t, n) for n, _ in GkcuTimings._fields_}

    def counters(self):
        c = GkcuCounters()
        _ck(lib().gkcu_get_counters(self._ctx, C.byref(c)))
        return {n: getattr(c, n) for n, _ in GkcuCounters._fields_}

    def set_tlas_refit(self, enabled, rebuild_every=0):
        _ck(lib().gkcu_set_tlas_refit(self._ctx, int(bool(enabled)), int(rebuild_every)))

    def tlas_stats(self):
        b, r = C.c_uint64(0), C.c_uint64(0)
        _ck(lib().gkcu_tlas_stats(self._ctx, C.byref(b), C.byref(r)))
        return {"builds": b.value, "refits": r.value}

    def set_skip_unchanged_nodes(self, enabled):
        _ck(lib().gkcu_set_skip_unchanged_nodes(self._ctx, int(bool(enabled))))

    def set_fast_stencil(self, enabled):
        _ck(lib().gkcu_set_fast_stencil(self._ctx, int(bool(enabled))))

    def set_kernel_timing(self, enabled):
        _ck(lib().gkcu_set_kernel_timing(self._ctx, int(bool(enabled))))

    def kernel_stats(self):
        k = GkcuKernelStats()
        _ck(lib().gkcu_get_kernel_stats(self._ctx, C.byref(k)))
        return {n: getattr(k, n) for n, _ in GkcuKernelStats._fields_}

    def readback_into(self, rt, ptr, nbytes):
        """Blocking device->host copy into caller-owned (e.g. pinned) host memory."""
        _ck(lib().gkcu_readback(self._ctx, rt, ptr, nbytes))

    def present_async(self, ptr, nbytes):
        """Copy of RT_DENOISED into (pinned) host memory behind the last frame, overlapping the next frame (gkcu_readback_denoised_async)."""
        _ck(lib().gkcu_readback_denoised_async(self._ctx, ptr, nbytes))

    def present_wait(self):
        _ck(lib().gkcu_readback_wait(self._ctx))

    EXT_OPAQUE_FD, EXT_CUDA_VMM_FD = 0, 1

    def bind_external_image(self, rt, handle_type, fd, size_bytes, offset_bytes=0, row_pitch_bytes=0):
        """Zero-copy present: RT_DENOISED lives in imported external memory (the engine's exportable render image, RenderImage.hpp:30-37)."""
        _ck(lib().gkcu_bind_external_image(self._ctx, rt, handle_type, fd, size_bytes, offset_bytes, row_pitch_bytes))

    def unbind_external_image(self, rt):
        _ck(lib().gkcu_unbind_external_image(self._ctx, rt))

    def bind_external_semaphore(self, handle_type, fd):
        _ck(lib().gkcu_bind_external_semaphore(self._ctx, handle_type, fd))

    def update_nodes_ptr(self, ptr, count):
        _ck(lib().gkcu_update_nodes(self._ctx, C.cast(ptr, C.POINTER(GkNodeProxy)), count))

    def stream(self):
        return lib().gkcu_stream(self._ctx)

    def trace(self, rays, any_hit=False):
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        hits = np.zeros(len(rays), dtype=HIT_DTYPE)
        _ck(lib().gkcu_trace_rays(self._ctx, rays.ctypes.data, len(rays), hits.ctypes.data, int(any_hit)))
        return hits

    # ---- validation hooks
    def debug_info(self):
        cap = 6 + 6 * 4096
        out = (C.c_uint32 * cap)()
        _ck(lib().gkcu_debug_info(self._ctx, out, cap))
        ext = 4 + 4 * out[3]
        models = [dict(node_base=out[4 + 4 * m], node_count=out[5 + 4 * m], tri_base=out[6 + 4 * m], prim_count=out[7 + 4 * m],
                       leaf_base=out[ext + 2 + 2 * m], leaf_count=out[ext + 3 + 2 * m])
                  for m in range(out[3])]
        return dict(tlas_nodes=out[0], tlas_levels=out[1], instances=out[2], models=models, node_format=out[ext], tlas_leaf_count=out[ext + 1])

    def debug_read(self, what, offset, count, dtype):
        arr = np.zeros(count, dtype=dtype)
        _ck(lib().gkcu_debug_read(self._ctx, what, offset, arr.ctypes.data, arr.nbytes))
        return arr

    def close(self):
        if self._ctx:
            lib().gkcu_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class RendererRegistry:
    """VulkanBaseRenderer::RegisterLogicRenderer / SwitchLogicRenderer (VulkanBaseRenderer.cpp:892-917) for the headless host."""

    def __init__(self, device=0):
        self.device = device
        self.logicRenderers_ = {}
        self.currentLogicRenderer_ = None

    def RegisterLogicRenderer(self, rtype):
        rtype = ERendererType(rtype)
        if rtype != ERendererType.ERT_CudaPathTracing:
            raise GkcuError(f"{rtype.name} is a Vulkan renderer of the reference and is not part of this build")
        self.logicRenderers_[rtype] = CudaPathTracingRenderer(self.device)
        self.currentLogicRenderer_ = rtype

    def SwitchLogicRenderer(self, rtype):
        self.currentLogicRenderer_ = ERendererType(rtype)

    def current(self):
        return self.logicRenderers_[self.currentLogicRenderer_]
