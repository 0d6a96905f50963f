"""gknextengine_b200 — B200-native (sm_100a) CUDA path-tracing hot path of gkNextEngine behind its renderer-switch interface.

csrc/   CUDA kernels + the C ABI (libgknext_cuda.so, include/gkcu.h)
host/   C++ host-side scene / UBO code the reference keeps on the CPU (libgknext_host.so)
"""
from . import host  # noqa: F401
from ._abi import RT, GkcuFrameParams, GkUniformBufferObject  # noqa: F401
from .renderer import CudaPathTracingRenderer, ERendererType, GkcuError, RendererRegistry  # noqa: F401
