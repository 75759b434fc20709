"""CPU suite for the post passes: the restatement (oracle/oracle.cpp) against the reference's
own PassMotionBlur / PassLightShafts compiled into oracle/_ref, and against the committed reference-generated fixtures.
RGBA8 outputs: everything is bit-exact."""
import os

import numpy as np
import pytest

import post_cases

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "golden_post_passes.npz")


@pytest.mark.parametrize("name", list(post_cases.blur_cases()))
def test_motion_blur_bit_exact(port, reference, name):
    make, p = post_cases.blur_cases()[name]
    ldr, depth, motion = make()
    a = port.pass_motion_blur(p, ldr, motion, depth)
    b = reference.pass_motion_blur(p, ldr, motion, depth)
    assert np.array_equal(a, b), f"{name}: {int(np.count_nonzero((a != b).any(axis=2)))} pixels differ"
    if p.enable:
        assert np.count_nonzero((a != ldr).any(axis=2)) > 0, "the blur must change something"
    else:
        assert np.array_equal(a, ldr)


@pytest.mark.parametrize("name", list(post_cases.shafts_cases()))
def test_light_shafts_bit_exact(port, reference, name):
    make, p, with_depth = post_cases.shafts_cases()[name]
    ldr, depth, _ = make()
    a = port.pass_light_shafts(p, ldr, depth if with_depth else None)
    b = reference.pass_light_shafts(p, ldr, depth if with_depth else None)
    assert np.array_equal(a, b), f"{name}: {int(np.count_nonzero((a != b).any(axis=2)))} pixels differ"
    if name in ("sun_behind", "sun_off_screen", "disabled"):
        assert np.array_equal(a, ldr), "the pass degenerates to a copy (pass_light_shafts.hpp:53-67, 96-108)"
    else:
        assert np.count_nonzero(a[..., :3] > ldr[..., :3]) > 0


def test_post_goldens(port):
    """Restatement vs the fixtures written by tests/golden/make_golden.py from the reference itself."""
    g = np.load(GOLDEN)
    for name, (make, p) in post_cases.blur_cases().items():
        ldr, depth, motion = make()
        assert np.array_equal(port.pass_motion_blur(p, ldr, motion, depth), g["blur_" + name]), name
    for name, (make, p, with_depth) in post_cases.shafts_cases().items():
        ldr, depth, _ = make()
        assert np.array_equal(port.pass_light_shafts(p, ldr, depth if with_depth else None), g["shafts_" + name]), name


def test_taa_restatement_properties(port):
    """The restatement against an independent numpy statement of pass_adapters.hpp:1471-1489 (the pin against the compiled
    reference is test_taa_restatement_equals_the_reference below)."""
    frames = post_cases.taa_frames()
    hist = np.zeros_like(frames[0])
    valid = False
    for f in frames:
        cur = f.copy()
        want_hist = hist.copy()
        if not valid:
            want, want_hist = cur.copy(), cur.copy()
        else:
            v = np.float32(1.0 - np.float32(0.12)) * cur[..., :3].astype(np.float32) + np.float32(0.12) * hist[..., :3].astype(np.float32)
            want = cur.copy()
            want[..., :3] = np.clip((v + np.float32(0.5)).astype(np.int32), 0, 255).astype(np.uint8)
            want_hist = want.copy()
        port.pass_taa(cur, hist, valid)
        valid = True
        assert np.array_equal(cur, want) and np.array_equal(hist, want_hist)


def test_taa_restatement_equals_the_reference(port):
    """PINNED: the reference's own PassTemporalAAAdapter (pipeline/pass_adapters.hpp:1402-1491) compiled against the JoltPhysics
    declaration shim (oracle/ref_taa_harness.cpp) vs the restatement, frame after frame with the history carried over -- including
    every (current, history) byte pair at least once (a 256 x 256 frame), so the rounding rule is covered exhaustively."""
    import os

    from oracle import bindings
    if not os.path.exists(bindings.REF_TAA_LIB) and not os.path.isdir("/root/reference"):
        pytest.skip("oracle/_ref/libshs_taa_ref.so not built and /root/reference absent")
    for frames in (post_cases.taa_frames(), post_cases.taa_frames(w=33, h=7, n=6, seed=3)):
        hp, hr = np.zeros_like(frames[0]), np.zeros_like(frames[0])
        valid = False
        for f in frames:
            a, b = f.copy(), f.copy()
            port.pass_taa(a, hp, valid)
            bindings.reference_pass_taa(b, hr, valid)
            valid = True
            assert np.array_equal(a, b) and np.array_equal(hp, hr)
    cur = np.zeros((256, 256, 4), np.uint8)
    cur[..., 0] = np.arange(256)[:, None]
    cur[..., 1] = np.arange(256)[:, None]
    cur[..., 2] = 255 - np.arange(256)[:, None]
    cur[..., 3] = 77
    hist = np.zeros((256, 256, 4), np.uint8)
    hist[..., 0] = np.arange(256)[None, :]
    hist[..., 1] = 255 - np.arange(256)[None, :]
    hist[..., 2] = np.arange(256)[None, :]
    hist[..., 3] = 255
    a, b, hp, hr = cur.copy(), cur.copy(), hist.copy(), hist.copy()
    port.pass_taa(a, hp, True)
    bindings.reference_pass_taa(b, hr, True)
    assert np.array_equal(a, b) and np.array_equal(hp, hr) and not np.array_equal(a, cur)
