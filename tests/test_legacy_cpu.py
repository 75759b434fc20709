"""The legacy tile-job rasterizer (BASELINE configs[0] "as shipped") on the CPU:
oracle/oracle_legacy.cpp against the reference's OWN demo sources (hello_pipeline_blinn_phong_shading.cpp + shs_renderer.hpp,
compiled where they lie by oracle/ref_legacy_harness.cpp) -- canvas and z-buffer bit for bit -- and the host helpers of the C-ABI
(shsb_legacy_camera / _world_matrix / _mvp, pure host code: callable without a device) against the reference's Camera3D / glm."""
import numpy as np
import pytest

import fuzz_cases
from leisure_software_renderer_b200 import renderer, scenes
from oracle.bindings import LegacyOracle


@pytest.fixture(scope="module")
def lport():
    return LegacyOracle("port")


@pytest.fixture(scope="module")
def lref():
    import os
    if not LegacyOracle.available("reference") and not os.path.isdir("/root/reference"):
        pytest.skip("oracle/_ref/libshs_legacy_ref.so not built and /root/reference absent")
    return LegacyOracle("reference")


def render(o, W, H, tile_w, tile_h, cam, light, objs, angles, helper=None):
    helper = helper or o
    view, proj = helper.camera(cam, *angles)
    canvas = np.zeros((H, W, 4), np.uint8)
    canvas[..., 3] = 255                                             # Canvas clears to opaque black
    canvas[..., 0] = 17                                              # ... a marker: uncovered pixels must keep their content
    z = np.full((H, W), np.finfo(np.float32).max, np.float32)        # ZBuffer::clear
    for pos, nrm, margs, color in objs:
        model = helper.world_matrix(*margs)
        o.draw(pos, nrm, helper.mvp(proj, view, model), model, light, cam, color, canvas, z, tile_w, tile_h)
    return canvas, z


def c1_inputs():
    """configs[0] exactly as the demo sets it up (hello_pipeline_blinn_phong_shading.cpp:152-153, 384): Viewer (0, 5, -20), Suzanne
    at (0, 0, 10) scaled 4, colour (60, 100, 200), light direction normalize(-1, -0.4, 1); canvas 640 x 480 (the BASELINE size)."""
    m = scenes.load_suzanne()
    pos, nrm = m["positions"][m["indices"]], m["normals"][m["indices"]]
    light = np.array([-1.0, -0.4, 1.0], np.float32)
    light = light * np.float32(1.0 / np.sqrt(np.float32(light @ light)))
    return 640, 480, 80, 80, (0.0, 5.0, -20.0), tuple(float(v) for v in light), [(pos, nrm, ((0.0, 0.0, 10.0), (4.0, 4.0, 4.0), 0.0), (60, 100, 200, 255))], (0.0, 0.0)


def test_c1_as_shipped_bit_exact(lport, lref):
    a = render(lport, *c1_inputs())
    b = render(lref, *c1_inputs())
    assert np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32)), "z-buffer differs"
    assert np.array_equal(a[0], b[0]), f"canvas differs at {int(np.count_nonzero((a[0] != b[0]).any(axis=2)))} px"
    covered = a[1] < np.finfo(np.float32).max
    assert 3000 < int(covered.sum()) < 100000, "Suzanne fills a plausible part of the frame"
    # the canvas is y-flipped with respect to the z-buffer (Canvas::draw_pixel_screen_space)
    assert np.array_equal(covered[::-1], (a[0][..., :3] != np.array([17, 0, 0], np.uint8)).any(axis=2))


@pytest.mark.parametrize("seed", list(range(120)))
def test_fuzz_legacy_bit_exact(lport, lref, seed):
    args = fuzz_cases.legacy_draws(seed)
    a = render(lport, *args)
    b = render(lref, *args)
    assert np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32)), f"seed {seed}: z-buffer differs at {int(np.count_nonzero(a[1] != b[1]))} px"
    assert np.array_equal(a[0], b[0]), f"seed {seed}: canvas differs at {int(np.count_nonzero((a[0] != b[0]).any(axis=2)))} px"


def test_host_helpers_match_the_reference(lref):
    """shsb_legacy_camera / _world_matrix / _mvp (C-ABI, host only) == Camera3D::update, MonkeyObject::get_world_matrix, proj*view*model."""
    rng = np.random.default_rng(5)

    class Host:
        camera = staticmethod(renderer.legacy_camera)
        world_matrix = staticmethod(renderer.legacy_world_matrix)
        mvp = staticmethod(renderer.legacy_mvp)

    for i in range(200):
        pos = rng.uniform(-20, 20, 3)
        ang = (0.0, 0.0) if i == 0 else (float(rng.uniform(-180, 180)), float(rng.uniform(-85, 85)))
        v0, p0 = lref.camera(pos, *ang)
        v1, p1 = Host.camera(pos, *ang)
        assert np.array_equal(v0.view(np.uint32), v1.view(np.uint32)) and np.array_equal(p0.view(np.uint32), p1.view(np.uint32)), (i, ang)
        margs = (rng.uniform(-5, 5, 3), rng.uniform(0.2, 5, 3), float(rng.uniform(-360, 360)))
        m0, m1 = lref.world_matrix(*margs), Host.world_matrix(*margs)
        assert np.array_equal(m0.view(np.uint32), m1.view(np.uint32)), i
        assert np.array_equal(lref.mvp(p0, v0, m0).view(np.uint32), Host.mvp(p0, v0, m0).view(np.uint32)), i


def test_legacy_goldens(lport):
    """The restatement against the committed fixture written by tests/golden/make_golden_legacy.py from the reference's own code."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import make_golden_legacy as mg
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_legacy.npz"))
    for name, make in mg.CASES.items():
        canvas, z = render(lport, *make())
        assert np.array_equal(z.view(np.uint32), g[name + "_z"].view(np.uint32)), name
        assert np.array_equal(canvas, g[name + "_canvas"]), name
