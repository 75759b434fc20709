"""GPU parity (`-m gpu`) of the legacy tile-job rasterizer behind the C-ABI (shsb_legacy_draw_blinn_phong, csrc/legacy.cu; BASELINE
configs[0] "as shipped") against the CPU oracle (oracle/oracle_legacy.cpp, which tests/test_legacy_cpu.py
pins bit for bit against the reference's own compiled demo sources).  Gates: coverage and the z-buffer bit-exact (every operation that
selects a pixel or a depth is the reference's, unfused), canvas <= 1 LSB per channel (the specular powf is CUDA's, not glibc's)."""
import numpy as np
import pytest

import fuzz_cases
from leisure_software_renderer_b200 import capi, renderer
from oracle.bindings import LegacyOracle
from test_legacy_cpu import c1_inputs, render

pytestmark = pytest.mark.gpu

FLT_MAX = np.finfo(np.float32).max


def gpu_render(gpu, W, H, tile_w, tile_h, cam, light, objs, angles, indexed_mesh=None):
    view, proj = renderer.legacy_camera(cam, *angles)
    canvas = np.zeros((H, W, 4), np.uint8)
    canvas[..., 3] = 255
    canvas[..., 0] = 17
    c_rt = gpu.rt_create(capi.RT_COLOR_LDR, W, H)
    z_rt = gpu.rt_create(capi.RT_SHADOW, W, H)
    meshes = []
    try:
        gpu.rt_upload(c_rt, capi.PLANE_COLOR, canvas)
        gpu.rt_clear(z_rt, capi.PLANE_DEPTH, np.float32(FLT_MAX))              # ZBuffer::clear
        for pos, nrm, margs, color in objs:
            model = renderer.legacy_world_matrix(*margs)
            u = capi.LegacyUniforms(renderer.legacy_mvp(proj, view, model), model, light, cam, color, tile_w, tile_h)
            m = indexed_mesh if indexed_mesh is not None else gpu.mesh_upload(pos, nrm, None, None)
            gpu.legacy_draw_blinn_phong(m, u, c_rt, z_rt)
        return gpu.rt_download(c_rt), gpu.rt_download(z_rt, capi.PLANE_DEPTH)
    finally:
        gpu.rt_destroy(c_rt)
        gpu.rt_destroy(z_rt)


def check(g, c, name):
    assert np.array_equal(g[1].view(np.uint32), c[1].view(np.uint32)), f"{name}: z-buffer differs at {int(np.count_nonzero(g[1].view(np.uint32) != c[1].view(np.uint32)))} px"
    d = np.abs(g[0].astype(np.int32) - c[0].astype(np.int32))
    assert int(d.max()) <= 1, f"{name}: canvas differs by {int(d.max())} LSB"
    return int(np.count_nonzero(d))


def test_c1_as_shipped(gpu):
    """configs[0] exactly as the demo sets it up, through the soup AND through the indexed Suzanne mesh (expanded on the device)."""
    from leisure_software_renderer_b200 import scenes
    port = LegacyOracle("port")
    args = c1_inputs()
    c = render(port, *args)
    check(gpu_render(gpu, *args), c, "c1 soup")
    m = scenes.load_suzanne()
    h = gpu.mesh_upload(m["positions"], m["normals"], m["uvs"], m["indices"])
    check(gpu_render(gpu, *args, indexed_mesh=h), c, "c1 indexed")
    assert int((c[1] < FLT_MAX).sum()) > 3000


@pytest.mark.parametrize("seed", list(range(40)))
def test_fuzz_legacy_parity(gpu, seed):
    port = LegacyOracle("port")
    args = fuzz_cases.legacy_draws(seed)
    check(gpu_render(gpu, *args), render(port, *args), f"seed {seed}")


def test_legacy_error_behaviour(gpu):
    import ctypes as C
    u = capi.LegacyUniforms(np.eye(4, dtype=np.float32).ravel(), np.eye(4, dtype=np.float32).ravel())
    c_rt = gpu.rt_create(capi.RT_COLOR_LDR, 32, 16)
    z_small = gpu.rt_create(capi.RT_SHADOW, 16, 16)
    z_ok = gpu.rt_create(capi.RT_SHADOW, 32, 16)
    hdr = gpu.rt_create(capi.RT_COLOR_HDR, 32, 16)
    tri = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)
    m = gpu.mesh_upload(tri, tri, None, None)
    m_no_normals = gpu.mesh_upload(tri, None, None, None)
    try:
        f = gpu.lib.shsb_legacy_draw_blinn_phong
        assert f(gpu.h, 9999, C.byref(u), c_rt, z_small) == 2          # unknown mesh
        assert f(gpu.h, m, C.byref(u), c_rt, z_small) == 7             # z-buffer of another size
        assert f(gpu.h, m, C.byref(u), c_rt, hdr) == 2                 # a target without a depth plane
        assert f(gpu.h, m, C.byref(u), hdr, z_small) == 2              # the canvas must be RGBA8
        assert f(gpu.h, m_no_normals, C.byref(u), c_rt, z_ok) == 1     # the legacy vertex shader needs a normal per position
        assert f(gpu.h, m, None, c_rt, z_ok) == 1
        assert f(gpu.h, m, C.byref(u), c_rt, z_ok) == 0
    finally:
        for rt in (c_rt, z_small, z_ok, hdr):
            gpu.rt_destroy(rt)
