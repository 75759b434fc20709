"""The legacy soft-shadow demo (config-3 flavour) on the CPU: the restatement (second half of
oracle/oracle_legacy.cpp) against the reference's OWN source, hello_shadow_mapping_soft.cpp, compiled where it lies by
oracle/ref_legacy2_harness.cpp -- shadow map, canvas and z-buffer bit for bit.  (This is the checker the CUDA path of the row, csrc/legacy2.cu, is compared with:
tests/test_zz_gpu_legacy2.py.)"""
import os

import numpy as np
import pytest

import fuzz_cases
from oracle.bindings import L2Uniforms, Legacy2Oracle

FLT_MAX = np.finfo(np.float32).max


@pytest.fixture(scope="module")
def l2port():
    return Legacy2Oracle("port")


@pytest.fixture(scope="module")
def l2ref():
    if not Legacy2Oracle.available("reference") and not os.path.isdir("/root/reference"):
        pytest.skip("oracle/_ref/libshs_legacy2_ref.so not built and /root/reference absent")
    return Legacy2Oracle("reference")


def render(o, sc, with_shadow=True):
    f32 = sc["f32"]
    shadow = np.full((sc["sm"], sc["sm"]), FLT_MAX, np.float32)
    for pos, nrm, uv, model, color, use_tex in sc["objs"]:
        o.shadow_draw(pos, f32(model), sc["light_vp"], shadow, *sc["tile"])
    canvas = np.zeros((sc["H"], sc["W"], 4), np.uint8)
    canvas[...] = (20, 20, 25, 255)                                   # the demo's clear colour (:1088)
    z = np.full((sc["H"], sc["W"]), FLT_MAX, np.float32)
    for pos, nrm, uv, model, color, use_tex in sc["objs"]:
        u = L2Uniforms()
        mv = sc["view"] @ model
        for name, m in (("mvp", sc["proj"] @ mv), ("model", model), ("mv", mv)):
            getattr(u, name)[:] = list(f32(m))
        nm = np.linalg.inv(model[:3, :3]).T
        u.normal_mat[:] = list(np.ascontiguousarray(np.asarray(nm, np.float32).T).reshape(9))
        u.light_vp[:] = list(sc["light_vp"])
        u.light_dir_world[:] = list(sc["light_dir"])
        u.camera_pos[:] = list(sc["cam"])
        u.base_color[:] = list(color)
        u.use_texture = int(use_tex)
        o.camera_draw(pos, nrm, uv, u, canvas, z, texture=sc["texture"], shadow=shadow if with_shadow else None, tile_w=sc["tile"][0], tile_h=sc["tile"][1])
    return shadow, canvas, z


@pytest.mark.parametrize("seed", list(range(60)))
def test_fuzz_legacy2_bit_exact(l2port, l2ref, seed):
    sc = fuzz_cases.legacy2_scene(seed)
    a = render(l2port, sc, with_shadow=seed % 6 != 5)
    b = render(l2ref, sc, with_shadow=seed % 6 != 5)
    assert np.array_equal(a[0].view(np.uint32), b[0].view(np.uint32)), f"seed {seed}: shadow map differs at {int(np.count_nonzero(a[0] != b[0]))} texels"
    assert np.array_equal(a[2].view(np.uint32), b[2].view(np.uint32)), f"seed {seed}: z-buffer differs at {int(np.count_nonzero(a[2] != b[2]))} px"
    assert np.array_equal(a[1], b[1]), f"seed {seed}: canvas differs at {int(np.count_nonzero((a[1] != b[1]).any(axis=2)))} px"


def test_fuzz_legacy2_scenes_are_not_trivial(l2port):
    drawn = shadowed = 0
    for seed in range(12):
        sc = fuzz_cases.legacy2_scene(seed)
        lit = render(l2port, sc, with_shadow=False)
        sh = render(l2port, sc, with_shadow=True)
        drawn += int((lit[2] < FLT_MAX).sum() > 200)
        shadowed += int(np.count_nonzero((lit[1] != sh[1]).any(axis=2)) > 20)
    assert drawn >= 10 and shadowed >= 6, (drawn, shadowed)
