"""GPU parity (`-m gpu`) of the legacy render-target demos behind the C-ABI (shsb_legacy2_shadow_draw, shsb_legacy2_draw_softshadow,
shsb_legacy3_draw_pbr; csrc/legacy2.cu) against the CPU oracle (oracle/oracle_legacy.cpp, pinned
bit for bit against the reference's own compiled hello_shadow_mapping_soft.cpp / hello_pbr.cpp by tests/test_legacy2_cpu.py and
tests/test_legacy3_cpu.py).  Gates: shadow map, z-buffer and velocity buffer bit-exact; canvas <= 1 LSB per channel (powf is CUDA's,
not glibc's) at EVERY pixel: the soft-shadow demo's PCSS kernel rotation reads the host libm's sinf / cosf of its 2^24 possible angles
from a table (api.cu: ensure_l2_rotation), so no tap moves across a texel boundary any more (round 1 allowed two pixels per scene).
(The file sorts last on purpose: these kernels were written after the round's GPU budget was spent and have only been checked
through the CPU emulation of their device functions; a failure here must not hide the rest of the GPU suite under `-x`.)"""
import numpy as np
import pytest

import fuzz_cases
import test_legacy2_cpu as t2
import test_legacy3_cpu as t3
from leisure_software_renderer_b200 import capi
from oracle.bindings import Legacy2Oracle, Legacy3Oracle

pytestmark = pytest.mark.gpu

FLT_MAX = np.finfo(np.float32).max


def fill_uniforms(sc, model, color, use_tex, tex_h, prev=False):
    f32 = sc["f32"]
    u = capi.Legacy2Uniforms()
    mv = sc["view"] @ model
    mats = [("mvp", sc["proj"] @ mv), ("model", model), ("mv", mv)]
    if prev:
        mats.append(("prev_mvp", sc["proj"] @ sc["prev_view"] @ model))
    for name, m in mats:
        getattr(u, name)[:] = list(f32(m))
    nm = np.linalg.inv(model[:3, :3]).T
    u.normal_mat[:] = list(np.ascontiguousarray(np.asarray(nm, np.float32).T).reshape(9))
    u.light_vp[:] = list(sc["light_vp"])
    u.light_dir_world[:] = list(sc["light_dir"])
    u.camera_pos[:] = list(sc["cam"])
    u.base_color[:] = list(color)
    u.use_texture = int(use_tex)
    u.albedo = tex_h
    u.job_tile_w, u.job_tile_h = sc["tile"]
    return u


def gpu_render(gpu, sc, with_shadow=True, pbr=False):
    W, H, sm = sc["W"], sc["H"], sc["sm"]
    canvas = np.zeros((H, W, 4), np.uint8)
    canvas[...] = (20, 20, 25, 255)
    c_rt = gpu.rt_create(capi.RT_COLOR_LDR, W, H)
    z_rt = gpu.rt_create(capi.RT_DEPTH_MOTION if pbr else capi.RT_SHADOW, W, H)
    s_rt = gpu.rt_create(capi.RT_SHADOW, sm, sm)
    tex = gpu.texture_upload(sc["texture"])
    ibl = gpu.legacy3_ibl_upload(sc["irradiance"], sc["prefiltered"]) if pbr and sc["irradiance"] is not None else 0
    try:
        gpu.rt_upload(c_rt, capi.PLANE_COLOR, canvas)
        gpu.rt_clear(z_rt, capi.PLANE_DEPTH, np.float32(FLT_MAX))
        gpu.rt_clear(s_rt, capi.PLANE_DEPTH, np.float32(FLT_MAX))
        if pbr:
            gpu.rt_upload(z_rt, capi.PLANE_MOTION, np.zeros((H, W, 2), np.float32))
        meshes = [gpu.mesh_upload(pos, nrm, uv, None) for pos, nrm, uv, *_ in sc["objs"]]
        for m, (pos, nrm, uv, model, color, use_tex) in zip(meshes, sc["objs"]):
            gpu.legacy2_shadow_draw(m, sc["f32"](model), sc["light_vp"], s_rt, *sc["tile"])
        for k, (m, (pos, nrm, uv, model, color, use_tex)) in enumerate(zip(meshes, sc["objs"])):
            u = fill_uniforms(sc, model, color, use_tex, tex, prev=pbr)
            if pbr:
                u.metallic, u.roughness, u.ao = sc["pbr"][k]
                u.ibl_diffuse_intensity, u.ibl_specular_intensity, u.ibl_reflection_strength = sc["ibl_k"]
                gpu.legacy3_draw_pbr(m, u, s_rt if with_shadow else 0, ibl, c_rt, z_rt)
            else:
                gpu.legacy2_draw_softshadow(m, u, s_rt if with_shadow else 0, c_rt, z_rt)
        out = [gpu.rt_download(s_rt, capi.PLANE_DEPTH), gpu.rt_download(c_rt), gpu.rt_download(z_rt, capi.PLANE_DEPTH)]
        if pbr:
            out.append(gpu.rt_download(z_rt, capi.PLANE_MOTION))
        return out
    finally:
        for rt in (c_rt, z_rt, s_rt):
            gpu.rt_destroy(rt)
        if ibl:
            gpu.legacy3_ibl_destroy(ibl)


def check(g, c, name, loose_pixels=0):
    for k, what in ((0, "shadow map"), (2, "z-buffer")) + (((3, "velocity"),) if len(g) > 3 else ()):
        a, b = g[k].view(np.uint32), c[k].view(np.uint32)
        assert np.array_equal(a, b), f"{name}: {what} differs at {int(np.count_nonzero(a != b))} of {a.size} words"
    d = np.abs(g[1].astype(np.int32) - c[1].astype(np.int32)).max(axis=2)
    assert int(np.count_nonzero(d > 1)) <= loose_pixels, f"{name}: canvas differs by more than 1 LSB at {int(np.count_nonzero(d > 1))} px (max {int(d.max())})"


@pytest.mark.parametrize("seed", list(range(24)))
def test_fuzz_legacy2_softshadow_parity(gpu, seed):
    sc = fuzz_cases.legacy2_scene(seed)
    ws = seed % 6 != 5
    check(gpu_render(gpu, sc, with_shadow=ws), t2.render(Legacy2Oracle("port"), sc, with_shadow=ws), f"L2 seed {seed}")


@pytest.mark.parametrize("seed", list(range(24)))
def test_fuzz_legacy3_pbr_parity(gpu, seed):
    sc = fuzz_cases.legacy3_scene(seed)
    ws = seed % 6 != 5
    check(gpu_render(gpu, sc, with_shadow=ws, pbr=True), t3.render(Legacy3Oracle("port"), sc, with_shadow=ws), f"L3 seed {seed}")


def test_legacy2_indexed_mesh_equals_soup(gpu):
    """The mesh handle's indices are expanded on the device like ModelGeometry's loader does on the host."""
    from leisure_software_renderer_b200 import scenes
    m = scenes.load_suzanne()
    sc = fuzz_cases.legacy2_scene(3)
    model = np.eye(4)
    sc["objs"] = [(m["positions"][m["indices"]], m["normals"][m["indices"]], m["uvs"][m["indices"]], model, (200, 120, 90, 255), True)]
    soup = gpu_render(gpu, sc)
    s_rt, c_rt, z_rt = gpu.rt_create(capi.RT_SHADOW, sc["sm"], sc["sm"]), gpu.rt_create(capi.RT_COLOR_LDR, sc["W"], sc["H"]), gpu.rt_create(capi.RT_SHADOW, sc["W"], sc["H"])
    try:
        canvas = np.zeros((sc["H"], sc["W"], 4), np.uint8)
        canvas[...] = (20, 20, 25, 255)
        gpu.rt_upload(c_rt, capi.PLANE_COLOR, canvas)
        gpu.rt_clear(z_rt, capi.PLANE_DEPTH, np.float32(FLT_MAX))
        gpu.rt_clear(s_rt, capi.PLANE_DEPTH, np.float32(FLT_MAX))
        h = gpu.mesh_upload(m["positions"], m["normals"], m["uvs"], m["indices"])
        gpu.legacy2_shadow_draw(h, sc["f32"](model), sc["light_vp"], s_rt, *sc["tile"])
        gpu.legacy2_draw_softshadow(h, fill_uniforms(sc, model, (200, 120, 90, 255), True, gpu.texture_upload(sc["texture"])), s_rt, c_rt, z_rt)
        assert np.array_equal(gpu.rt_download(s_rt, capi.PLANE_DEPTH), soup[0])
        assert np.array_equal(gpu.rt_download(z_rt, capi.PLANE_DEPTH), soup[2])
        assert np.array_equal(gpu.rt_download(c_rt), soup[1])
    finally:
        for rt in (s_rt, c_rt, z_rt):
            gpu.rt_destroy(rt)


def test_legacy2_error_behaviour(gpu):
    import ctypes as C
    u = capi.Legacy2Uniforms()
    c_rt = gpu.rt_create(capi.RT_COLOR_LDR, 32, 16)
    z_small = gpu.rt_create(capi.RT_SHADOW, 16, 16)
    z_ok = gpu.rt_create(capi.RT_SHADOW, 32, 16)
    dm_ok = gpu.rt_create(capi.RT_DEPTH_MOTION, 32, 16)
    hdr = gpu.rt_create(capi.RT_COLOR_HDR, 32, 16)
    tri = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)
    m = gpu.mesh_upload(tri, tri, None, None)
    m_no_normals = gpu.mesh_upload(tri, None, None, None)
    eye = np.eye(4, dtype=np.float32).ravel()
    try:
        f = gpu.lib.shsb_legacy2_draw_softshadow
        assert f(gpu.h, 9999, C.byref(u), 0, c_rt, z_ok) == 2            # unknown mesh
        assert f(gpu.h, m, C.byref(u), 0, c_rt, z_small) == 7            # z-buffer of another size
        assert f(gpu.h, m, C.byref(u), 0, c_rt, hdr) == 2                # a target without a depth plane
        assert f(gpu.h, m, C.byref(u), 0, hdr, z_ok) == 2                # the canvas must be RGBA8
        assert f(gpu.h, m, C.byref(u), hdr, c_rt, z_ok) == 2             # the shadow map must be a depth target
        assert f(gpu.h, m, C.byref(u), z_ok, c_rt, z_ok) == 1            # shadow map == z-buffer
        assert f(gpu.h, m_no_normals, C.byref(u), 0, c_rt, z_ok) == 1    # the vertex shader needs a normal per position
        assert f(gpu.h, m, None, 0, c_rt, z_ok) == 1
        assert f(gpu.h, m, C.byref(u), 0, c_rt, z_ok) == 0
        g = gpu.lib.shsb_legacy3_draw_pbr
        assert g(gpu.h, m, C.byref(u), 0, 0, c_rt, z_ok) == 2            # the PBR demo's target carries a velocity plane
        assert g(gpu.h, m, C.byref(u), 0, 77, c_rt, dm_ok) == 2          # unknown IBL handle
        assert g(gpu.h, m, C.byref(u), 0, 0, c_rt, dm_ok) == 0
        s = gpu.lib.shsb_legacy2_shadow_draw
        assert s(gpu.h, m_no_normals, capi.fptr(eye), capi.fptr(eye), 0, 0, z_small) == 0   # the shadow pass reads positions only
        assert s(gpu.h, m, capi.fptr(eye), capi.fptr(eye), 0, 0, c_rt) == 2
        assert s(gpu.h, m, None, capi.fptr(eye), 0, 0, z_small) == 1
        assert s(gpu.h, m, capi.fptr(eye), capi.fptr(eye), -1, 0, z_small) == 1
    finally:
        for rt in (c_rt, z_small, z_ok, dm_ok, hdr):
            gpu.rt_destroy(rt)


@pytest.mark.parametrize("name", ["legacy2_drop_in_test", "legacy3_drop_in_test"])
def test_legacy2_drop_in_parity_on_gpu(name):
    """The demos' own draw_triangle_tile_* loops on the CPU vs shs::b200::legacy2::Renderer fed with the same Uniforms objects
    (tests/cpp/legacy2_drop_in_test.cpp): shadow map, z-buffer and velocity bit-equal, canvas within 1 LSB."""
    import os
    import subprocess
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cpp", "_build", name)
    if not os.path.exists(path):
        pytest.skip(f"tests/cpp/_build/{name} was not built (needs /root/reference at build time)")
    r = subprocess.run([path], capture_output=True, text=True, timeout=300)
    print(r.stdout, r.stderr)
    assert r.returncode == 0, r.stdout + r.stderr
