"""The restatement against the committed reference-generated fixtures (runs anywhere, no /root/reference needed)."""
import os
import sys

import numpy as np
import pytest

import harness

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
import make_golden  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("name", list(make_golden.GOLDEN))
def test_port_matches_reference_golden(port, name):
    make, kw = make_golden.GOLDEN[name]
    g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    f = harness.cpu_forward(port, make(), aov=False, **kw)
    assert np.array_equal(f.hdr.view(np.uint32), g["hdr"].view(np.uint32))
    assert np.array_equal(f.ldr, g["ldr"])
    if "depth" in g:
        assert np.array_equal(f.depth.view(np.uint32), g["depth"].view(np.uint32))
    if "shadow" in g:
        assert np.array_equal(f.shadow.view(np.uint32), g["shadow"].view(np.uint32))
        assert np.array_equal(f.lvp.view(np.uint32), g["lvp"].view(np.uint32))
    assert [f.stats[k] for k in ("tri_input", "tri_after_clip", "tri_raster")] == list(g["stats"])


@pytest.mark.parametrize("name", list(make_golden.GOLDEN_MOTION))
def test_port_matches_reference_motion_golden(port, name):
    prev, cur = make_golden.GOLDEN_MOTION[name]()
    g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    f = harness.cpu_forward(port, cur, aov=False, motion=True, prev_models=prev.models(port))
    assert np.array_equal(f.motion.view(np.uint32), g["motion"].view(np.uint32))
    assert np.array_equal(f.hdr.view(np.uint32), g["hdr"].view(np.uint32)) and np.array_equal(f.depth.view(np.uint32), g["depth"].view(np.uint32))


def test_light_lists_golden(port):
    from leisure_software_renderer_b200 import scenes
    sd = scenes.scene_small(w=320, h=200, lights=64)
    g = np.load(os.path.join(GOLDEN_DIR, "golden_light_lists_320x200_port.npz"))
    counts, indices = port.light_cull(sd.lights, sd.viewproj, sd.w, sd.h, 16, 128)
    assert np.array_equal(counts, g["counts"]) and np.array_equal(indices, g["indices"])
