"""GPU parity (`-m gpu`) of the depth-range / clustered light-bin builders and the per-tile depth reduce behind the C-ABI
(shsb_light_cull_ex, shsb_tile_depth_range) against the CPU oracle and the committed fixture: bin lists bit-exact.
(The reference's own headers for these need JoltPhysics: parity is against the restatement, 'unpinned'.)"""
import os

import numpy as np
import pytest

import harness
import test_light_bins_cpu as lb
from leisure_software_renderer_b200 import capi, scenes

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "golden_light_bins_port.npz")


def _render_depth(gpu, sd):
    g = harness.GpuScene(gpu, sd)
    fp = capi.FrameParams.from_buffer_copy(sd.fp)
    fp.light_culling = 0
    gpu.history_reset()
    gpu.pass_pbr_forward(sd.scene, fp, g.hdr, g.dm)
    return g, gpu.rt_download(g.dm, capi.PLANE_DEPTH)


def test_bins_small_scene_all_modes(gpu, port):
    sd = lb.bins_scene()
    gold = np.load(GOLDEN)
    g, depth = _render_depth(gpu, sd)
    try:
        lo, hi = gpu.tile_depth_range(g.dm, 16)
        olo, ohi = port.tile_depth_range(depth, 16, sd.zn, sd.zf)
        assert np.array_equal(lo.view(np.uint32), olo.view(np.uint32)) and np.array_equal(hi.view(np.uint32), ohi.view(np.uint32))
        assert np.array_equal(lo, gold["range_min"]) and np.array_equal(hi, gold["range_max"]), "depth (and so its ranges) equals the CPU frame's bit for bit"
        for name, d in lb.descs(sd, mx=32).items():
            rng = (lo, hi) if name == "view_depth" else ((gold["range01_min"], gold["range01_max"]) if name == "depth01" else (None, None))
            c, i = gpu.light_cull_ex(d, *rng)
            oc, oi = port.light_cull_ex(sd.lights, d, *rng)
            assert np.array_equal(c, oc), f"{name}: counts differ in {int(np.count_nonzero(c != oc))} bins"
            keep = np.arange(32)[None, :] < np.minimum(oc, 32)[:, None]
            assert np.array_equal(i[keep], oi[keep]), f"{name}: indices differ"
            assert np.array_equal(c, gold[name + "_counts"]) and np.array_equal(i[keep], gold[name + "_indices"][keep]), f"{name}: fixture"
        # device-resident ranges (no host arrays): same lists as with the downloaded ranges
        d = lb.descs(sd)["view_depth"]
        c_dev, i_dev = gpu.light_cull_ex(d)
        c_host, i_host = gpu.light_cull_ex(d, lo, hi)
        assert np.array_equal(c_dev, c_host) and np.array_equal(i_dev, i_host)
    finally:
        g.release()


def test_forward_plus_with_depth_range_lists_is_unchanged(gpu):
    """Lights removed by the depth range cannot reach any fragment of the tile (cull sphere radius >= range), so the lit
    frame is identical with the tighter lists (no list is saturated in this scene)."""
    sd = lb.bins_scene()
    g, _ = _render_depth(gpu, sd)
    try:
        fp = capi.FrameParams.from_buffer_copy(sd.fp)
        fp.light_culling = 1
        gpu.light_cull(sd.viewproj, sd.w, sd.h, 16, 128)
        c_plain, _ = gpu.light_lists_download()
        assert int(c_plain.max()) <= 128
        gpu.pass_pbr_forward(sd.scene, fp, g.hdr, g.dm)
        a = gpu.rt_download(g.hdr)
        gpu.tile_depth_range(g.dm, 16)
        c_tight, _ = gpu.light_cull_ex(lb.descs(sd)["view_depth"])
        assert int(c_tight.sum()) < int(c_plain.sum())
        gpu.pass_pbr_forward(sd.scene, fp, g.hdr, g.dm)
        b = gpu.rt_download(g.hdr)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    finally:
        g.release()


def test_bins_error_behaviour(gpu):
    import ctypes as C
    sd = lb.bins_scene()
    d = lb.descs(sd)["view_depth"]
    d.viewport_w = 999   # no device ranges for this viewport
    assert gpu.lib.shsb_light_cull_ex(gpu.h, C.byref(d), None, None) == 1
    d = lb.descs(sd)["clustered"]
    d.depth_slices = 0
    assert gpu.lib.shsb_light_cull_ex(gpu.h, C.byref(d), None, None) == 1
    d.mode = 9
    assert gpu.lib.shsb_light_cull_ex(gpu.h, C.byref(d), None, None) == 1
    assert gpu.lib.shsb_tile_depth_range(gpu.h, 9999, 16) == 2


def test_bins_full_1080p(gpu, port):
    """BASELINE configs[1]: 1920x1080, 1024 lights -- depth reduce + view-depth lists and 8-slice clusters, whole frame vs oracle."""
    sd = scenes.scene_c2()
    g, depth = _render_depth(gpu, sd)
    try:
        lo, hi = gpu.tile_depth_range(g.dm, 16)
        olo, ohi = port.tile_depth_range(depth, 16, sd.zn, sd.zf)
        assert np.array_equal(lo, olo) and np.array_equal(hi, ohi)
        for mode, kw in ((capi.LIGHT_CULL_TILED_VIEW_DEPTH, {}), (capi.LIGHT_CULL_CLUSTERED, {"depth_slices": 8})):
            d = capi.LightCullDesc(sd.viewproj, sd.w, sd.h, mode, 16, 64, z_near=sd.zn, z_far=sd.zf, **kw)
            rng = (lo, hi) if mode == capi.LIGHT_CULL_TILED_VIEW_DEPTH else (None, None)
            c, i = gpu.light_cull_ex(d, *rng)
            oc, oi = port.light_cull_ex(sd.lights, d, *rng)
            assert np.array_equal(c, oc), f"mode {mode}: counts differ in {int(np.count_nonzero(c != oc))} bins"
            keep = np.arange(64)[None, :] < np.minimum(oc, 64)[:, None]
            assert np.array_equal(i[keep], oi[keep])
    finally:
        g.release()


def test_bins_with_every_light_type(gpu, port):
    """The reference-packed mixed light set (rect / tube area lights carry OBB- and capsule-derived cull boxes, Jolt-bounded points a
    sqrt(3)-inflated sphere): all four bin builders against the oracle."""
    import cases
    sd = cases._mixed(w=208, h=120)
    g, depth = _render_depth(gpu, sd)
    try:
        lo, hi = gpu.tile_depth_range(g.dm, 16)
        for name, d in lb.descs(sd, mx=64).items():
            rng = (lo, hi) if name in ("view_depth",) else ((np.clip(lo / np.float32(sd.zf), 0, 1), np.clip(hi / np.float32(sd.zf), 0, 1)) if name == "depth01" else (None, None))
            c, i = gpu.light_cull_ex(d, *rng)
            oc, oi = port.light_cull_ex(sd.lights, d, *rng)
            assert np.array_equal(c, oc), f"{name}: counts differ in {int(np.count_nonzero(c != oc))} bins"
            keep = np.arange(64)[None, :] < np.minimum(oc, 64)[:, None]
            assert np.array_equal(i[keep], oi[keep]), f"{name}: indices differ"
            assert int(oc.sum()) > 0
    finally:
        g.release()


@pytest.mark.parametrize("seed", list(range(25)))
def test_depth_reduce_ndc01_equals_the_pinned_restatement(gpu, port, seed):
    """shsb_tile_depth_range_ndc01 (the reference's fp_stress_depth_reduce.comp as a device path) against the restatement that
    tests/test_light_bins_cpu.py pins to the shader's own text: per-tile (min, max) bit for bit, for three tile sizes, on depth planes
    with cleared / out-of-range / negative texels, empty tiles, ragged sizes and the shader's near / far clamps."""
    d, zn, zf = lb.ndc01_depth_buffers(seed)
    h, w = d.shape
    rt = gpu.rt_create(capi.RT_SHADOW if seed % 2 else capi.RT_DEPTH_MOTION, w, h)
    try:
        gpu.rt_upload(rt, capi.PLANE_DEPTH, d)
        for ts in (16, 8, 32):
            lo, hi = gpu.tile_depth_range_ndc01(rt, ts, zn, zf)
            olo, ohi = port.tile_depth_range_ndc01(d, ts, zn, zf)
            assert np.array_equal(lo.view(np.uint32), olo.view(np.uint32)) and np.array_equal(hi.view(np.uint32), ohi.view(np.uint32)), (seed, ts)
    finally:
        gpu.rt_destroy(rt)
    assert gpu.lib.shsb_tile_depth_range_ndc01(gpu.h, 9999, 16, 0.1, 100.0) == 2 and gpu.lib.shsb_tile_depth_range_ndc01(gpu.h, 1, 0, 0.1, 100.0) == 1
