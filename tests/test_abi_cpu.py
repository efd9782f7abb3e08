"""CPU-side checks of the boundary: the C ABI library loads and exports every symbol include/clothgpu.h
declares, the PODs have the layout the header says, and the library fails loudly without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import clothgpu
from clothgpu import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "clothgpu.h")).read()
    return sorted(set(re.findall(r"CLOTHGPU_API[^;(]*?\b(clothgpu_\w+)\s*\(", hdr)))


def test_header_and_binding_agree_on_the_symbol_list():
    assert _declared_symbols() == sorted(clothgpu.EXPORTS)


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build()
    L = C.CDLL(clothgpu.LIB_PATH)
    for name in _declared_symbols():
        assert hasattr(L, name), name


def test_pod_layouts():
    assert C.sizeof(_abi.Desc) == 16 * 4
    assert C.sizeof(_abi.Frame) == 128
    assert _abi.Frame.mouse_x.offset == 24 and _abi.Frame.dt.offset == 96 and _abi.Frame.tear_length.offset == 112
    assert C.sizeof(_abi.Counters) == 64
    assert C.sizeof(_abi.Info) == 56


def test_frame_default_matches_the_reference_defaults():
    import __graft_entry__ as ge
    ge.build()
    f = _abi.Frame()
    assert clothgpu.lib().clothgpu_frame_default(C.byref(f)) == 0
    g = _abi.default_frame()
    assert bytes(f) == bytes(g)
    assert list(f.slider) == [2.0, 250.0, 30.0, float(np.float32(0.98)), 15.0]
    d = _abi.Desc()
    assert clothgpu.lib().clothgpu_desc_default(C.byref(d)) == 0
    assert bytes(d) == bytes(_abi.default_desc())


def test_no_cpu_fallback():
    """Without a CUDA device create() must fail with NODEVICE, never compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import __graft_entry__ as ge
    ge.build()
    with pytest.raises(clothgpu.ClothGpuError) as e:
        clothgpu.Cloth(_abi.default_desc())
    assert e.value.code == _abi.ERR_NODEVICE
    assert "no CPU fallback" in str(e.value)


def test_product_sources_never_reference_the_oracle():
    pkg = os.path.join(ROOT, "cloth-physics_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".cu", ".cuh", ".h", ".hpp", ".cpp", ".cc", ".py", ".go")):
                src = open(os.path.join(dirpath, fn), errors="ignore").read()
                assert "oracle" not in src.lower(), os.path.join(dirpath, fn)
