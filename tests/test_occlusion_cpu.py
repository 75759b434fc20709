"""Software occlusion (geometry/culling_software.hpp:253-333) on the CPU box:
  * the restatement (oracle/oracle_scene_cull.cpp: shso_software_occlusion) equals the reference's OWN header compiled against the
    JoltPhysics declaration shim (oracle/ref_occlusion_harness.cpp) -- occluded flags, the ordered visible list, the four
    CullingStats counters and the occlusion depth buffer, bit for bit;
  * the DEVICE functions (csrc/scene_cull_core.cuh) compiled by g++ and driven like the kernel drives them -- minimum on the
    depth's bit pattern, texels and triangles visited in reverse -- equal the restatement bit for bit."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import fuzz_cases
from oracle import bindings

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(os.path.dirname(HERE), "leisure_software_renderer_b200", "csrc")


def emul():
    out, src, hdr = os.path.join(HERE, "cpp", "_build", "libscene_cull_emul.so"), os.path.join(HERE, "cpp", "scene_cull_emul.cpp"), os.path.join(CSRC, "scene_cull_core.cuh")
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-I" + CSRC, src, "-o", out], check=True)
    return bindings.SoftwareOcclusion(out, "shsemu_")


def same(a, b, what):
    for x, y, name in zip(a, b, ("occluded flags", "visible list", "counts", "depth buffer")):
        if name == "depth buffer":
            assert np.array_equal(np.asarray(x) + 0.0, np.asarray(y) + 0.0), f"{what}: {name} differs at {int(np.count_nonzero(x != y))} texels"   # -0 == +0
        else:
            assert np.array_equal(x, y), f"{what}: {name} differ: {x} vs {y}"


@pytest.mark.skipif(not bindings.SoftwareOcclusion.available(), reason="oracle/_ref/libshs_occlusion_ref.so not built and /root/reference absent")
@pytest.mark.parametrize("seed", list(range(60)))
def test_fuzz_restatement_equals_the_reference(seed):
    sc = fuzz_cases.occlusion_scene(seed)
    ref, port = bindings.SoftwareOcclusion("reference"), bindings.SoftwareOcclusion("port")
    for enable in (True, False):
        r, p = ref.run(sc, enable), port.run(sc, enable)
        same(p[:3] + ((p[3],) if enable else ()), r[:3] + ((r[3],) if enable else ()), f"seed {seed} enable {enable}")
        assert int(r[2][2]) + int(r[2][3]) == int(r[2][1])


@pytest.mark.parametrize("seed", list(range(60)))
def test_fuzz_device_functions_equal_the_restatement(seed):
    sc = fuzz_cases.occlusion_scene(seed)
    p, e = bindings.SoftwareOcclusion("port").run(sc), emul().run(sc)
    same(e, p, f"seed {seed}")


def test_the_fuzz_scenes_occlude_and_show():
    """The comparison is not vacuous: across the seeds objects are occluded, visible, skipped as stale, and depth gets written."""
    port = bindings.SoftwareOcclusion("port")
    occluded = visible = written = 0
    for seed in range(60):
        sc = fuzz_cases.occlusion_scene(seed)
        occ, vis, counts, depth = port.run(sc)
        occluded += int(occ.sum()); visible += len(vis); written += int(np.count_nonzero(depth < 1.0))
    assert occluded > 100 and visible > 500 and written > 10000, (occluded, visible, written)
