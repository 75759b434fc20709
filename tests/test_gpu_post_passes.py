"""GPU parity (`-m gpu`) of the post passes behind the C-ABI (shsb_pass_motion_blur / _light_shafts / _taa) against the
CPU oracle, the committed reference-generated fixtures, and -- at the 1080p size of BASELINE configs[1] -- against the
oracle on the full frame plus size-independent properties.  RGBA8 outputs: every comparison is bit-exact."""
import os

import numpy as np
import pytest

import post_cases
from leisure_software_renderer_b200 import capi

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "golden_post_passes.npz")


class Targets:
    def __init__(self, gpu, ldr, depth=None, motion=None, out_size=None):
        h, w = ldr.shape[:2]
        self.gpu = gpu
        self.src = gpu.rt_create(capi.RT_COLOR_LDR, w, h)
        ow, oh = out_size or (w, h)
        self.dst = gpu.rt_create(capi.RT_COLOR_LDR, ow, oh)
        self.dm = gpu.rt_create(capi.RT_DEPTH_MOTION, w, h, 0.1, 1000.0)
        gpu.rt_upload(self.src, capi.PLANE_COLOR, ldr)
        if depth is not None:
            gpu.rt_upload(self.dm, capi.PLANE_DEPTH, depth)
        if motion is not None:
            gpu.rt_upload(self.dm, capi.PLANE_MOTION, motion)

    def close(self):
        for rt in (self.src, self.dst, self.dm):
            self.gpu.rt_destroy(rt)


@pytest.mark.parametrize("name", list(post_cases.blur_cases()))
def test_motion_blur_parity(gpu, port, name):
    make, p = post_cases.blur_cases()[name]
    ldr, depth, motion = make()
    want = port.pass_motion_blur(p, ldr, motion, depth)
    assert np.array_equal(want, np.load(GOLDEN)["blur_" + name]), "oracle vs reference fixture"
    t = Targets(gpu, ldr, depth, motion)
    try:
        gpu.pass_motion_blur(p, t.src, t.dst, t.dm)
        got = gpu.rt_download(t.dst)
        assert np.array_equal(got, want), f"{name}: {int(np.count_nonzero((got != want).any(axis=2)))} pixels differ"
        gpu.pass_motion_blur(p, t.src, t.src, t.dm)   # in place: the reference's scratch path
        assert np.array_equal(gpu.rt_download(t.src), want)
    finally:
        t.close()


@pytest.mark.parametrize("name", list(post_cases.shafts_cases()))
def test_light_shafts_parity(gpu, port, name):
    make, p, with_depth = post_cases.shafts_cases()[name]
    ldr, depth, _ = make()
    want = port.pass_light_shafts(p, ldr, depth if with_depth else None)
    assert np.array_equal(want, np.load(GOLDEN)["shafts_" + name]), "oracle vs reference fixture"
    t = Targets(gpu, ldr, depth)
    try:
        gpu.pass_light_shafts(p, t.src, t.dst, t.dm if with_depth else 0)
        got = gpu.rt_download(t.dst)
        assert np.array_equal(got, want), f"{name}: {int(np.count_nonzero((got != want).any(axis=2)))} pixels differ"
        gpu.pass_light_shafts(p, t.src, t.src, t.dm if with_depth else 0)
        assert np.array_equal(gpu.rt_download(t.src), want)
    finally:
        t.close()


def test_post_passes_crop_to_common_size(gpu, port):
    """Targets of different sizes: the reference works on min(w), min(h) (pass_motion_blur.hpp:51-53, pass_light_shafts.hpp:72-73)."""
    ldr, depth, motion = post_cases.planes(64, 48, 9)
    t = Targets(gpu, ldr, depth, motion, out_size=(50, 40))
    try:
        sentinel = np.full((40, 50, 4), 7, dtype=np.uint8)
        p = capi.MotionBlurParams()
        gpu.rt_upload(t.dst, capi.PLANE_COLOR, sentinel)
        gpu.pass_motion_blur(p, t.src, t.dst, t.dm)
        # the oracle entry takes equally sized planes: crop the inputs the way the pass indexes them (same rows / columns, w = 50, h = 40)
        want = port.pass_motion_blur(p, ldr[:40, :50].copy(), motion[:40, :50].copy(), depth[:40, :50].copy())
        assert np.array_equal(gpu.rt_download(t.dst), want)
        # a depth_like target whose size differs from the common size is ignored by the shafts (pass_light_shafts.hpp:165)
        make, sp, _ = post_cases.shafts_cases()["sun_in_view"]
        gpu.pass_light_shafts(sp, t.src, t.dst, t.dm)
        want = port.pass_light_shafts(sp, ldr[:40, :50].copy(), None)
        assert np.array_equal(gpu.rt_download(t.dst), want)
    finally:
        t.close()


def test_taa_parity(gpu, port):
    frames = post_cases.taa_frames()
    h, w = frames[0].shape[:2]
    rt = gpu.rt_create(capi.RT_COLOR_LDR, w, h)
    try:
        for reset_at in (None, 2):
            gpu.taa_reset()
            hist = np.zeros_like(frames[0])
            valid = False
            for i, f in enumerate(frames):
                if reset_at == i:
                    gpu.taa_reset()          # PassTemporalAAAdapter::reset_history
                    valid = False
                cur = f.copy()
                port.pass_taa(cur, hist, valid)
                valid = True
                gpu.rt_upload(rt, capi.PLANE_COLOR, f)
                gpu.pass_taa(rt)
                assert np.array_equal(gpu.rt_download(rt), cur), f"frame {i}"
    finally:
        gpu.rt_destroy(rt)
        gpu.taa_reset()


def test_post_error_behaviour(gpu):
    lib = gpu.lib
    import ctypes as C
    p = capi.MotionBlurParams()
    assert lib.shsb_pass_motion_blur(gpu.h, C.byref(p), 9999, 9998, 9997) == 2     # SHSB_E_INVALID_HANDLE
    assert lib.shsb_pass_motion_blur(gpu.h, None, 1, 1, 1) == 1                     # SHSB_E_INVALID_ARGUMENT
    sp = capi.LightShaftsParams()
    assert lib.shsb_pass_light_shafts(gpu.h, C.byref(sp), 9999, 9998, 0) == 2
    assert lib.shsb_pass_taa(gpu.h, 9999) == 2


def test_post_passes_full_1080p(gpu, port):
    """BASELINE configs[1] frame size: the whole 1920x1080 frame against the oracle, plus determinism."""
    w, h = 1920, 1080
    ldr, depth, motion = post_cases.planes(w, h, 21, max_motion=60.0)
    t = Targets(gpu, ldr, depth, motion)
    try:
        p = capi.MotionBlurParams(samples=16, max_velocity_px=32.0)
        gpu.pass_motion_blur(p, t.src, t.dst, t.dm)
        a = gpu.rt_download(t.dst)
        assert np.array_equal(a, port.pass_motion_blur(p, ldr, motion, depth))
        gpu.pass_motion_blur(p, t.src, t.dst, t.dm)
        assert np.array_equal(a, gpu.rt_download(t.dst)), "deterministic"
        _, sp, _ = post_cases.shafts_cases()["sun_in_view"]
        sp = capi.LightShaftsParams(cam_viewproj=list(sp.cam_viewproj), cam_pos=list(sp.cam_pos), sun_dir_ws=list(sp.sun_dir_ws))
        gpu.pass_light_shafts(sp, t.src, t.dst, t.dm)
        b = gpu.rt_download(t.dst)
        assert np.array_equal(b, port.pass_light_shafts(sp, ldr, depth))
        assert np.count_nonzero(b[..., :3] > ldr[..., :3]) > 0
    finally:
        t.close()
