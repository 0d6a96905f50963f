import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# Several tests run 2-3 renderer contexts on ONE GPU as stand-ins for 2-3 GPUs, with device-side barriers between their streams: a kernel of one
# context waits for a kernel of another. Streams that share a hardware queue would serialise those kernels, so give the device more queues than
# the default 8 (must be set before the CUDA context exists). One process per GPU, as in production and bench.py, does not need it.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def _ensure_built():
    pkg = os.path.join(ROOT, "gknextengine_b200")
    need = [os.path.join(pkg, "libgknext_cuda.so"), os.path.join(pkg, "libgknext_host.so"), os.path.join(ROOT, "oracle", "liboracle.so")]
    if not all(os.path.exists(p) for p in need):
        import __graft_entry__
        __graft_entry__.build()


_ensure_built()


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def sky():
    from gknextengine_b200 import host
    return host.make_sky(256, 128)


@pytest.fixture()
def renderer():
    from gknextengine_b200 import CudaPathTracingRenderer
    r = CudaPathTracingRenderer(0)
    yield r
    r.close()
