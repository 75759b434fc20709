"""The legacy PBR / IBL demo (config-4 flavour) on the CPU: the restatement (third part of
oracle/oracle_legacy.cpp) against the reference's OWN source, hello_pbr.cpp with the library's shs/resources/ibl.hpp, compiled where
it lies by oracle/ref_legacy3_harness.cpp -- shadow map, z-buffer and velocity buffer bit for bit, canvas identical.  (This is the checker the CUDA path of the row,
csrc/legacy2.cu in MODE_PBR, is compared with: tests/test_zz_gpu_legacy2.py.)"""
import os

import numpy as np
import pytest

import fuzz_cases
from oracle.bindings import L3Uniforms, Legacy3Oracle

FLT_MAX = np.finfo(np.float32).max


@pytest.fixture(scope="module")
def l3port():
    return Legacy3Oracle("port")


@pytest.fixture(scope="module")
def l3ref():
    if not Legacy3Oracle.available("reference") and not os.path.isdir("/root/reference"):
        pytest.skip("oracle/_ref/libshs_legacy3_ref.so not built and /root/reference absent")
    return Legacy3Oracle("reference")


def render(o, sc, with_shadow=True):
    f32 = sc["f32"]
    shadow = np.full((sc["sm"], sc["sm"]), FLT_MAX, np.float32)
    for pos, nrm, uv, model, color, use_tex in sc["objs"]:
        o.shadow_draw(pos, f32(model), sc["light_vp"], shadow, *sc["tile"])
    canvas = np.zeros((sc["H"], sc["W"], 4), np.uint8)
    canvas[...] = (20, 20, 25, 255)
    z = np.full((sc["H"], sc["W"]), FLT_MAX, np.float32)
    vel = np.zeros((sc["H"], sc["W"], 2), np.float32)
    for (pos, nrm, uv, model, color, use_tex), (metallic, roughness, ao) in zip(sc["objs"], sc["pbr"]):
        u = L3Uniforms()
        mv = sc["view"] @ model
        for name, m in (("mvp", sc["proj"] @ mv), ("prev_mvp", sc["proj"] @ sc["prev_view"] @ model), ("model", model), ("mv", mv)):
            getattr(u, name)[:] = list(f32(m))
        nm = np.linalg.inv(model[:3, :3]).T
        u.normal_mat[:] = list(np.ascontiguousarray(np.asarray(nm, np.float32).T).reshape(9))
        u.light_vp[:] = list(sc["light_vp"])
        u.light_dir_world[:] = list(sc["light_dir"])
        u.camera_pos[:] = list(sc["cam"])
        u.base_color_srgb[:] = list(color)
        u.metallic, u.roughness, u.ao = metallic, roughness, ao
        u.use_texture = int(use_tex)
        u.ibl_diffuse_intensity, u.ibl_specular_intensity, u.ibl_reflection_strength = sc["ibl_k"]
        o.camera_draw(pos, nrm, uv, u, canvas, z, vel, texture=sc["texture"], shadow=shadow if with_shadow else None,
                      irradiance=sc["irradiance"], prefiltered=sc["prefiltered"], tile_w=sc["tile"][0], tile_h=sc["tile"][1])
    return shadow, canvas, z, vel


@pytest.mark.parametrize("seed", list(range(60)))
def test_fuzz_legacy3_bit_exact(l3port, l3ref, seed):
    sc = fuzz_cases.legacy3_scene(seed)
    a = render(l3port, sc, with_shadow=seed % 6 != 5)
    b = render(l3ref, sc, with_shadow=seed % 6 != 5)
    assert np.array_equal(a[0].view(np.uint32), b[0].view(np.uint32)), f"seed {seed}: shadow map differs at {int(np.count_nonzero(a[0] != b[0]))} texels"
    assert np.array_equal(a[2].view(np.uint32), b[2].view(np.uint32)), f"seed {seed}: z-buffer differs at {int(np.count_nonzero(a[2] != b[2]))} px"
    assert np.array_equal(a[3].view(np.uint32), b[3].view(np.uint32)), f"seed {seed}: velocity differs at {int(np.count_nonzero((a[3] != b[3]).any(axis=2)))} px"
    assert np.array_equal(a[1], b[1]), f"seed {seed}: canvas differs at {int(np.count_nonzero((a[1] != b[1]).any(axis=2)))} px"


def test_fuzz_legacy3_scenes_are_not_trivial(l3port):
    drawn = moving = ibl_lit = 0
    for seed in range(12):
        sc = fuzz_cases.legacy3_scene(seed)
        full = render(l3port, sc)
        drawn += int((full[2] < FLT_MAX).sum() > 200)
        moving += int(np.count_nonzero(full[3]) > 50)
        if sc["irradiance"] is not None:
            no_ibl = render(l3port, dict(sc, irradiance=None, prefiltered=None))
            ibl_lit += int(np.count_nonzero((no_ibl[1] != full[1]).any(axis=2)) > 50)
    assert drawn >= 10 and moving >= 7 and ibl_lit >= 7, (drawn, moving, ibl_lit)
