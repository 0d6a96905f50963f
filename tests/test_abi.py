"""The C-ABI library loads and exports every symbol include/gkcu.h declares; layouts match; no silent CPU fallback."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT, has_gpu
from gknextengine_b200 import _abi


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "gkcu.h")).read()
    return sorted(set(re.findall(r"\b(gkcu_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(os.path.join(ROOT, "gknextengine_b200", "libgknext_cuda.so"))
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/gkcu.h but not exported"


def test_struct_sizes_match_reference_layout():
    # SURVEY appendix A, measured from the reference's BasicTypes.slang compiled as C++
    assert C.sizeof(_abi.GkUniformBufferObject) == 784
    assert _abi.GkUniformBufferObject.TotalFrames.offset == 668
    assert _abi.GkUniformBufferObject.NumberOfBounces.offset == 680
    assert _abi.GkUniformBufferObject.HasSun.offset == 708
    assert _abi.GkUniformBufferObject.BFSize.offset == 756
    assert _abi.GkUniformBufferObject.SuperResolution.offset == 776
    assert C.sizeof(_abi.GkNodeProxy) == 208 and _abi.GkNodeProxy.combinedPrevTS.offset == 80 and _abi.GkNodeProxy.matId.offset == 144
    assert C.sizeof(_abi.GkModelData) == 64 and _abi.GkModelData.reorderOffset.offset == 56
    assert C.sizeof(_abi.GkMaterial) == 64 and _abi.GkMaterial.MaterialModel.offset == 36
    assert C.sizeof(_abi.GkGPUVertex) == 24


def test_reference_layout_check_binary():
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_layout_check")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref not built (reference not mounted)")
    import subprocess
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout


@pytest.mark.skipif(has_gpu(), reason="checks the no-GPU failure mode")
def test_no_gpu_fails_loudly_instead_of_falling_back():
    from gknextengine_b200 import CudaPathTracingRenderer, GkcuError
    with pytest.raises(GkcuError):
        CudaPathTracingRenderer(0)


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "gknextengine_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                text = open(os.path.join(dp, f), errors="replace").read()
                assert "oracle_lib" not in text and "liboracle" not in text and "gk_oracle" not in text, f"{f} references the oracle"
