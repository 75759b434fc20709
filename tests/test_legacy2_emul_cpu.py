"""The DEVICE functions of the legacy render-target demos (rows L2 / L3; leisure_software_renderer_b200/csrc/legacy2_core.cuh, the
header legacy2.cu's kernels are made of) compiled by g++ and walked per pixel the way the raster kernel walks them
(tests/cpp/legacy2_emul.cpp), against the pinned restatement (oracle/oracle_legacy.cpp).  What this checks without a GPU: the set-up
records, the near-plane clip / fan slots, the job tiles' clamped integer boxes, and the order-independent reformulation of the
demos' serial loops (running minimum depth + last shadeable prefix minimum) -- z-buffer, shadow map, velocity and canvas bit for bit
with libm's sinf / cosf; with the device's flavour of the PCSS rotation (double sin / cos rounded to float) the canvas may differ
where a rotated tap sits on a texel boundary: counted and bounded here."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import fuzz_cases
import test_legacy2_cpu as t2
import test_legacy3_cpu as t3
from oracle.bindings import Legacy2Oracle, Legacy3Oracle

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "..", "leisure_software_renderer_b200", "csrc")


def emul_lib(flavour):
    out = os.path.join(HERE, "cpp", "_build", f"liblegacy2_emul_{flavour}.so")
    src = os.path.join(HERE, "cpp", "legacy2_emul.cpp")
    hdr = os.path.join(CSRC, "legacy2_core.cuh")
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        cmd = ["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-I" + CSRC, src, "-o", out]
        if flavour == "libm":
            cmd.insert(1, "-DEMUL_LIBM")
        subprocess.run(cmd, check=True)
    return C.CDLL(out)


class Emul2(Legacy2Oracle):
    def __init__(self, flavour):
        self.kind, self.lib, self.prefix = "emul", emul_lib(flavour), "shsemu_l2_"


class Emul3(Legacy3Oracle):
    def __init__(self, flavour):
        self.kind, self.lib, self.prefix = "emul", emul_lib(flavour), "shsemu_l3_"


@pytest.fixture(scope="module")
def ports():
    return Legacy2Oracle("port"), Legacy3Oracle("port")


def same(a, b):
    return all(np.array_equal(x.view(np.uint8), y.view(np.uint8)) for x, y in zip(a, b))


@pytest.mark.parametrize("seed", list(range(40)))
def test_l2_device_functions_equal_the_oracle(ports, seed):
    sc = fuzz_cases.legacy2_scene(seed)
    a = t2.render(Emul2("libm"), sc, with_shadow=seed % 6 != 5)
    b = t2.render(ports[0], sc, with_shadow=seed % 6 != 5)
    assert np.array_equal(a[0].view(np.uint32), b[0].view(np.uint32)), "shadow map"
    assert np.array_equal(a[2].view(np.uint32), b[2].view(np.uint32)), "z-buffer"
    assert np.array_equal(a[1], b[1]), f"canvas differs at {int(np.count_nonzero((a[1] != b[1]).any(axis=2)))} px"


@pytest.mark.parametrize("seed", list(range(40)))
def test_l3_device_functions_equal_the_oracle(ports, seed):
    sc = fuzz_cases.legacy3_scene(seed)
    a = t3.render(Emul3("libm"), sc, with_shadow=seed % 6 != 5)
    b = t3.render(ports[1], sc, with_shadow=seed % 6 != 5)
    for k, name in ((0, "shadow map"), (2, "z-buffer"), (3, "velocity")):
        assert np.array_equal(a[k].view(np.uint32), b[k].view(np.uint32)), name
    assert np.array_equal(a[1], b[1]), f"canvas differs at {int(np.count_nonzero((a[1] != b[1]).any(axis=2)))} px"


def test_l2_device_sincos_flavour_is_bounded(ports):
    """The PCSS rotation with double-precision sin / cos rounded to float instead of libm's sinf / cosf: depth planes are untouched,
    and over 24 scenes at most a handful of pixels may move (none has so far)."""
    moved = total = 0
    for seed in range(24):
        sc = fuzz_cases.legacy2_scene(seed)
        a = t2.render(Emul2("dev"), sc)
        b = t2.render(ports[0], sc)
        assert np.array_equal(a[0].view(np.uint32), b[0].view(np.uint32)) and np.array_equal(a[2].view(np.uint32), b[2].view(np.uint32))
        moved += int(np.count_nonzero((a[1] != b[1]).any(axis=2)))
        total += a[1].shape[0] * a[1].shape[1]
    assert moved <= 4, (moved, total)
