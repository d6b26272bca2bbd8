"""CPU tests of the boundary: libvvgpu.so loads, exports every symbol include/vvgpu.h declares, and the
product path fails loudly (no CPU fallback) when there is no CUDA device."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from vvflow_b200 import build, capi
    build.build()
    L = capi.load()
    header = open(os.path.join(ROOT, "include", "vvgpu.h")).read()
    declared = sorted(set(re.findall(r"\b(vvgpu_[a-z0-9_]+)\s*\(", header)))
    assert declared, "no declarations found"
    for name in declared:
        assert hasattr(L, name), f"libvvgpu.so does not export {name}"
    assert sorted(capi.SYMBOLS) == declared


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from vvflow_b200 import capi, vvhd
    with pytest.raises(capi.VVGpuError):
        capi.Context(0)
    with pytest.raises(capi.VVGpuError):
        vvhd.Space()


def test_product_does_not_import_oracle():
    """only tests/, __graft_entry__.smoke() and bench.py may touch oracle/"""
    pkg = os.path.join(ROOT, "vvflow_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("from_oracle", "").replace("oracle/pyref", ""), f
