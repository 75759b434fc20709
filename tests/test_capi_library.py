"""The C-ABI library builds for sm_100a, loads, and exports every symbol include/shsb.h declares.
No compute calls here (no GPU on the build box)."""
import ctypes as C
import os
import re

import pytest

from leisure_software_renderer_b200 import build, capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "shsb.h")).read()
    return sorted(set(re.findall(r"SHSB_API\s+[\w\s\*]+?\b(shsb_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    build.build()
    lib = C.CDLL(capi.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    bound = capi.load_library()
    assert set(names) == set(bound._shsb_symbols), set(names) ^ set(bound._shsb_symbols)


def test_struct_sizes_match_header():
    # sizes implied by include/shsb.h on LP64
    assert C.sizeof(capi.Stats) == 40
    assert C.sizeof(capi.RasterCfg) == 16
    assert C.sizeof(capi.Uniforms) == 3 * 64 + 4 * 16 + 8 * 4
    assert C.sizeof(capi.Transform) == 36
    assert C.sizeof(capi.RenderItem) == 36 + 8 + 12 + 12 + 12
    assert C.sizeof(capi.Scene) == 64 + 16 + 16 + 16 + 8
    assert C.sizeof(capi.FrameParams) == 64


def test_no_cpu_fallback_when_no_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    lib = capi.load_library()
    h = C.c_void_p()
    assert lib.shsb_context_create(0, C.byref(h)) == 3  # SHSB_E_NO_DEVICE
    from leisure_software_renderer_b200.renderer import Context
    with pytest.raises(capi.ShsbError):
        Context(0)


def test_built_for_sm_100a_only():
    out = os.popen(f"cuobjdump -lelf {capi.LIB_PATH} 2>/dev/null").read()
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs
