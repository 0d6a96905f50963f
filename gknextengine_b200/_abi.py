"""ctypes mirrors of include/gkcu.h (which mirrors the reference's BasicTypes.slang:22-291)."""
import ctypes as C

import numpy as np


class GkVec4(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float), ("w", C.c_float)]


class GkMat4(C.Structure):
    _fields_ = [("m", (C.c_float * 4) * 4)]


class GkUniformBufferObject(C.Structure):
    _fields_ = (
        [(n, GkMat4) for n in ("ModelView", "Projection", "ModelViewInverse", "ProjectionInverse", "ViewProjection",
                               "PrevViewProjection", "ViewProjectionUnJit", "PrevViewProjectionUnJit")]
        + [(n, GkVec4) for n in ("ViewportRect", "SunDirection", "SunColor", "BackGroundColor")]
        + [("SunViewProjection", GkMat4)]
        + [(n, C.c_float) for n in ("Aperture", "FocusDistance", "SkyRotation", "HeatmapScale", "PaperWhiteNit", "SkyIntensity")]
        + [(n, C.c_uint32) for n in ("SkyIdx", "TotalFrames", "MaxNumberOfBounces", "NumberOfSamples", "NumberOfBounces",
                                     "RandomSeed", "LightCount", "HasSky", "ShowHeatmap", "UseCheckerBoard", "TemporalFrames",
                                     "HasSun", "HDR", "AdaptiveSample")]
        + [("AdaptiveVariance", C.c_float)]
        + [(n, C.c_uint32) for n in ("AdaptiveSteps", "TAA", "SelectedId", "ShowEdge", "ProgressiveRender")]
        + [(n, C.c_float) for n in ("BFSigma", "BFSigmaLum", "BFSigmaNormal")]
        + [(n, C.c_uint32) for n in ("BFSize", "FastGather", "FastInterpole", "DebugDraw_Lighting", "DisableSpatialReuse",
                                     "SuperResolution", "_pad0")]
    )


class GkNodeProxy(C.Structure):
    _fields_ = [("instanceId", C.c_uint32), ("modelId", C.c_uint32), ("visible", C.c_uint32), ("nort", C.c_uint32),
                ("worldTS", GkMat4), ("combinedPrevTS", GkMat4), ("matId", C.c_uint32 * 16)]


class GkModelData(C.Structure):
    _fields_ = [("indexOffset", C.c_uint32), ("indexCount", C.c_uint32), ("vertexOffset", C.c_uint32), ("vertexCount", C.c_uint32),
                ("localAabbMin", GkVec4), ("localAabbMax", GkVec4), ("modelType", C.c_uint32), ("voxelDataIdx", C.c_uint32),
                ("reorderOffset", C.c_uint32), ("reserved2", C.c_uint32)]


class GkMaterial(C.Structure):
    _fields_ = [("Diffuse", GkVec4), ("DiffuseTextureId", C.c_int32), ("MRATextureId", C.c_int32), ("NormalTextureId", C.c_int32),
                ("Fuzziness", C.c_float), ("RefractionIndex", C.c_float), ("MaterialModel", C.c_uint32), ("Metalness", C.c_float),
                ("RefractionIndex2", C.c_float), ("NormalTextureScale", C.c_float), ("Reserverd2", C.c_float),
                ("Reserverd3", C.c_float), ("Reserverd4", C.c_float)]


class GkGPUVertex(C.Structure):
    _fields_ = [(n, C.c_uint16) for n in ("posx", "posy", "posz", "texcoordx", "normalx", "normaly", "normalz", "texcoordy",
                                          "tangentx", "tangenty", "tangentz", "tangentw")]


class GkcuSceneDesc(C.Structure):
    _fields_ = [("vertices", C.POINTER(GkGPUVertex)), ("vertex_count", C.c_uint32), ("simple_vertices", C.POINTER(C.c_uint16)),
                ("indices", C.POINTER(C.c_uint32)), ("index_count", C.c_uint32), ("offsets", C.POINTER(GkModelData)),
                ("offset_count", C.c_uint32), ("model_count", C.c_uint32)]


class GkcuFrameParams(C.Structure):
    _fields_ = [("progressive", C.c_uint32), ("temporal_frames", C.c_int32), ("fast_reproject", C.c_uint32), ("spp_rank", C.c_uint32),
                ("spp_world", C.c_uint32), ("tracer", C.c_uint32), ("skip_post", C.c_uint32), ("primary_visibility", C.c_uint32)]


class GkGPUDrivenStat(C.Structure):  # BasicTypes.slang:13-20
    _fields_ = [(n, C.c_uint32) for n in ("ProcessedCount", "VisibleCount", "CulledCount", "TriangleCount", "CulledTriangleCount", "pad0", "pad1", "pad2")]


DRAW_CMD_DTYPE = np.dtype([("indexCount", "<u4"), ("instanceCount", "<u4"), ("firstIndex", "<u4"), ("vertexOffset", "<i4"), ("firstInstance", "<u4")])  # VkDrawIndexedIndirectCommand


class GkcuTimings(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("tlas_update_ms", "rt_pass_ms", "reproject_pass_ms", "compose_pass_ms", "copy_pass_ms",
                                         "frame_ms", "blas_build_ms", "blas_refit_ms")]


class GkcuCounters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("primary_rays", "closest_hit_rays", "any_hit_rays", "wavefront_iterations", "kernel_launches")]


class GkcuKernelStats(C.Structure):
    _fields_ = [("primary_ms", C.c_float), ("trace_ms", C.c_float), ("shade_ms", C.c_float), ("tail_ms", C.c_float),
                ("trace_launches", C.c_uint32), ("shade_launches", C.c_uint32), ("trace_rays", C.c_uint64), ("tail_rays", C.c_uint64)]


class GkcuRay(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("ox", "oy", "oz", "tmin", "dx", "dy", "dz", "tmax")]


class GkcuHit(C.Structure):
    _fields_ = [("t", C.c_float), ("u", C.c_float), ("v", C.c_float), ("inst", C.c_uint32), ("prim", C.c_uint32)]


assert C.sizeof(GkUniformBufferObject) == 784
assert C.sizeof(GkNodeProxy) == 208 and C.sizeof(GkModelData) == 64 and C.sizeof(GkMaterial) == 64 and C.sizeof(GkGPUVertex) == 24

RAY_DTYPE = np.dtype([(n, "<f4") for n in ("ox", "oy", "oz", "tmin", "dx", "dy", "dz", "tmax")])
HIT_DTYPE = np.dtype([("t", "<f4"), ("u", "<f4"), ("v", "<f4"), ("inst", "<u4"), ("prim", "<u4")])

# storage images: reference formats (src/Rendering/VulkanBaseRenderer.cpp:439-457) -> (numpy dtype, channels)
RT = dict(ACCUMLATE_DIFFUSE=0, SINGLE_DIFFUSE=1, MINIGBUFFER=2, OBJEDCTID_0=3, OBJEDCTID_1=4, MOTIONVECTOR=5, ALBEDO=6, NORMAL=7,
          DENOISED=9, PREV_DEPTHBUFFER=10, ACCUMLATE_SPECULAR=11, SINGLE_SPECULAR=12, ACCUMLATE_ALBEDO=13, MOTIONMOMENT=14,
          SINGLE_PREV_DIFFUSE=20, SINGLE_PREV_SPECULAR=21, SINGLE_PREV_ALBEDO=22, EXT_PRIMITIVEID=64)
RT_FORMAT = {
    0: (np.float16, 4), 1: (np.float16, 4), 2: (np.uint32, 1), 3: (np.uint32, 1), 4: (np.uint32, 1), 5: (np.float32, 2), 6: (np.float16, 4),
    7: (np.float16, 4), 9: (np.float16, 4), 10: (np.float32, 1), 11: (np.float16, 4), 12: (np.float16, 4), 13: (np.float16, 4),
    14: (np.uint16, 1), 20: (np.float16, 4), 21: (np.float16, 4), 22: (np.float16, 4), 64: (np.uint32, 1),
}
OBJECTID_MISS = 65535


class GkLightObject(C.Structure):  # BasicTypes.slang:171-183
    _fields_ = [("p0", GkVec4), ("p1", GkVec4), ("p3", GkVec4), ("normal_area", GkVec4), ("lightMatIdx", C.c_uint32), ("reserved1", C.c_uint32),
                ("reserved2", C.c_uint32), ("reserved3", C.c_uint32)]


class GkSphericalHarmonics(C.Structure):  # BasicTypes.slang:164-169
    _fields_ = [("coefficients", (C.c_float * 9) * 3), ("padding", C.c_float)]


GK_CUBE_SIZE_XY, GK_CUBE_SIZE_Z, GK_PAGE_COUNT = 192, 48, 64
GK_CUBE_COUNT = GK_CUBE_SIZE_XY * GK_CUBE_SIZE_XY * GK_CUBE_SIZE_Z
