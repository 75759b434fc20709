"""Flat-shaded mesh draws, the consumer of the per-object light selections (last link of the demo's chain) on the
CPU box:
  * the restatement (oracle/oracle_flat_draw.cpp: shso_flat_draw) equals the reference's OWN text -- sw_render/debug_draw.hpp,
    the four ILightModel::sample of lighting/light_runtime.hpp and draw_mesh_multi_light_transformed cut out of
    exp-plumbing/hello_light_types_culling_sw.cpp, compiled by oracle/ref_flat_draw_harness.cpp -- canvas byte for byte, depth
    buffer bit for bit, both draw kinds;
  * the DEVICE functions (csrc/flat_draw_core.cuh) compiled by g++ and driven like the kernels drive them -- a minimum on
    (depth bits, running triangle number) per texel, triangles and texels visited in reverse -- equal the restatement: depth bit for
    bit, canvas within 1 LSB (std::pow / std::cos are narrowed from double there; observed: 0 differing channels)."""
import os
import subprocess

import numpy as np
import pytest

import fuzz_cases
from oracle import bindings

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(os.path.dirname(HERE), "leisure_software_renderer_b200", "csrc")
SEEDS = list(range(60))


def emul():
    out, src = os.path.join(HERE, "cpp", "_build", "libflat_draw_emul.so"), os.path.join(HERE, "cpp", "flat_draw_emul.cpp")
    hdrs = [os.path.join(CSRC, "flat_draw_core.cuh"), os.path.join(CSRC, "scene_cull_core.cuh")]
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(p) for p in [src] + hdrs):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-I" + CSRC, src, "-o", out], check=True)
    return bindings.FlatDraw(out, "shsemu_")


def same_depth(a, b, what):
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), f"{what}: depth differs at {int(np.count_nonzero(a.view(np.uint32) != b.view(np.uint32)))} texels"


@pytest.mark.skipif(not bindings.FlatDraw.available(), reason="oracle/_ref/libshs_flat_draw_ref.so not built and /root/reference absent")
@pytest.mark.parametrize("seed", SEEDS)
def test_fuzz_restatement_equals_the_reference(seed):
    sc = fuzz_cases.flat_draw_scene(seed)
    ref, port = bindings.FlatDraw("reference"), bindings.FlatDraw("port")
    for mode in (0, 1):
        (rc, rd), (pc, pd) = ref.run(sc, mode), port.run(sc, mode)
        same_depth(pd, rd, f"seed {seed} mode {mode}")
        assert np.array_equal(pc, rc), f"seed {seed} mode {mode}: canvas differs at {int(np.count_nonzero((pc != rc).any(axis=2)))} texels"


@pytest.mark.parametrize("seed", SEEDS)
def test_fuzz_device_functions_equal_the_restatement(seed):
    sc = fuzz_cases.flat_draw_scene(seed, dangling=seed % 2 == 1)
    port, dev = bindings.FlatDraw("port"), emul()
    for mode in (0, 1):
        (pc, pd), (ec, ed) = port.run(sc, mode), dev.run(sc, mode)
        same_depth(ed, pd, f"seed {seed} mode {mode}")
        diff = np.abs(ec.astype(np.int16) - pc.astype(np.int16))
        assert diff.max() <= 1, f"seed {seed} mode {mode}: canvas differs by {int(diff.max())}"
        assert int(np.count_nonzero(diff)) == 0, f"seed {seed} mode {mode}: {int(np.count_nonzero(diff))} channels differ by 1 LSB (double-rounding of pow / cos)"


def test_the_fuzz_scenes_draw_light_and_tie():
    """The comparison is not vacuous: texels get drawn, lights of all four models contribute, stale selection entries are met,
    equal-depth repeats are resolved to the earlier draw, and the pre-filled depth block rejects fragments."""
    port = bindings.FlatDraw("port")
    drawn = lit = 0
    types = set()
    for seed in SEEDS:
        sc = fuzz_cases.flat_draw_scene(seed)
        c1, d1 = port.run(sc, 1)
        drawn += int(np.count_nonzero(d1 != sc["depth"]))
        dark = dict(sc)
        dark["sel_counts"] = np.zeros_like(sc["sel_counts"])
        c0, _ = port.run(dark, 1)
        lit += int(np.count_nonzero((c0 != c1).any(axis=2)))
        types |= set(int(t) for t in sc["lights"]["light_type"])
    assert drawn > 200000 and lit > 20000 and types == {1, 2, 3, 4}, (drawn, lit, types)


def test_equal_depths_keep_the_earlier_draw():
    sc = fuzz_cases.flat_draw_scene(3, size=(64, 48))
    two = {k: (np.concatenate([v[:1], v[:1]]) if k in ("draw_mesh", "models", "base", "sel_counts", "sel_idx") else v) for k, v in sc.items()}
    two["base"] = np.array([[1.0, 0.0, 0.0], [0.0, 0.0, 1.0]], np.float32)
    one = {k: (v[:1] if k in ("draw_mesh", "models", "base", "sel_counts", "sel_idx") else v) for k, v in two.items()}
    for lib in (bindings.FlatDraw("port"), emul()):
        (c2, d2), (c1, d1) = lib.run(two, 0), lib.run(one, 0)
        assert np.array_equal(c2, c1) and np.array_equal(d2, d1)
