"""GPU differential fuzzing (`-m gpu`): the CUDA path through the C-ABI against the CPU oracle on the seeded random scenes of
scenes.scene_fuzz (the same generator tests/test_fuzz_cpu.py uses to pin the oracle against the compiled reference, bit for bit,
on the CPU).  Gates as everywhere: coverage masks, triangle ids, fragment counts, statistics and light lists bit-exact, depth and
shadow depth <= 1 ULP, LDR <= 1 LSB, HDR PSNR >= 60 dB."""
import numpy as np
import pytest

import fuzz_cases
import harness
from leisure_software_renderer_b200 import scenes

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", list(range(24)))
def test_fuzz_forward_parity(gpu, port, seed):
    sd = scenes.scene_fuzz(seed, zero_normals=False)
    shadow = bool(sd.fp.shadow_enable)
    depth = seed % 7 != 3
    g = harness.gpu_forward(gpu, sd, shadow=shadow, depth=depth)
    c = harness.cpu_forward(port, sd, shadow=shadow, depth=depth)
    harness.assert_frame_parity(g, c, depth=depth, name=sd.name)
    if shadow:
        assert np.array_equal(g.lvp.view(np.uint32), c.lvp.view(np.uint32)), f"{sd.name}: light camera differs"
        assert int(harness.ulp_diff(g.shadow, c.shadow).max()) <= 1, f"{sd.name}: shadow map differs by more than 1 ULP"


@pytest.mark.parametrize("seed", list(range(100, 110)))
def test_fuzz_forward_plus_parity(gpu, port, seed):
    sd = scenes.scene_fuzz(seed, lights=True, zero_normals=False)
    g = harness.gpu_forward(gpu, sd, forward_plus=True, fused=bool(seed & 1))
    c = harness.cpu_forward(port, sd, forward_plus=True)
    harness.assert_frame_parity(g, c, name=sd.name)


@pytest.mark.parametrize("seed", list(range(10)))
def test_fuzz_motion_vectors_parity(gpu, port, seed):
    """Two consecutive random frames: the context's history comes from rendering the first one; motion vectors bit-exact."""
    prev, cur = fuzz_cases.motion_pair(seed, zero_normals=False)
    g = harness.gpu_forward(gpu, cur, prev_scene=prev, motion=True)
    c = harness.cpu_forward(port, cur, prev_models=prev.models(port), motion=True)
    harness.assert_frame_parity(g, c, name=cur.name)
    assert g.motion is not None and np.array_equal(g.motion.view(np.uint32), c.motion.view(np.uint32))


@pytest.mark.parametrize("seed", list(range(16)))
def test_fuzz_post_passes_parity(gpu, port, seed):
    """PassMotionBlur / PassLightShafts with random (also out-of-range) parameters: RGBA8, bit-exact."""
    from test_gpu_post_passes import Targets
    ldr, depth, motion, p, q, with_depth = fuzz_cases.post_inputs(seed)
    t = Targets(gpu, ldr, depth, motion)
    try:
        want = port.pass_motion_blur(p, ldr, motion, depth)
        gpu.pass_motion_blur(p, t.src, t.dst, t.dm)
        got = gpu.rt_download(t.dst)
        assert np.array_equal(got, want), f"blur seed {seed}: {int(np.count_nonzero((got != want).any(axis=2)))} pixels differ"
        want = port.pass_light_shafts(q, ldr, depth if with_depth else None)
        gpu.pass_light_shafts(q, t.src, t.dst, t.dm if with_depth else 0)
        got = gpu.rt_download(t.dst)
        assert np.array_equal(got, want), f"shafts seed {seed}: {int(np.count_nonzero((got != want).any(axis=2)))} pixels differ"
    finally:
        t.close()


@pytest.mark.parametrize("seed", list(range(24)))
def test_fuzz_light_bins_parity(gpu, port, seed):
    """Tile light lists are a bit-exact gate of the north_star: random light sets / cameras / tile sizes / caps through the plain
    builder and all four modes of shsb_light_cull_ex (host-supplied depth ranges incl. zero-thickness cells), counts and lists
    equal to the oracle's."""
    lights, descs = fuzz_cases.light_bins(seed)
    gpu.lights_upload(lights.view(np.uint8))
    d0 = descs[0][1]
    vp = np.array(list(d0.view_proj), np.float32)
    gpu.light_cull(vp, d0.viewport_w, d0.viewport_h, d0.tile_size, d0.max_per_bin)
    c, i = gpu.light_lists_download()
    oc, oi = port.light_cull(lights, vp, d0.viewport_w, d0.viewport_h, d0.tile_size, d0.max_per_bin)
    keep = np.arange(d0.max_per_bin)[None, :] < np.minimum(oc, d0.max_per_bin)[:, None]
    assert np.array_equal(c, oc) and np.array_equal(i[keep], oi[keep]), f"seed {seed}: plain tiled lists differ"
    for name, d, lo, hi in descs:
        c, i = gpu.light_cull_ex(d, lo, hi)
        oc, oi = port.light_cull_ex(lights, d, lo, hi)
        assert np.array_equal(c, oc), f"seed {seed} {name}: counts differ in {int(np.count_nonzero(c != oc))} of {c.size} bins"
        keep = np.arange(d.max_per_bin)[None, :] < np.minimum(oc, d.max_per_bin)[:, None]
        assert np.array_equal(i[keep], oi[keep]), f"seed {seed} {name}: lists differ"


@pytest.mark.parametrize("seed", list(range(10)))
def test_fuzz_light_lists_far_camera_high_resolution(gpu, port, seed):
    """The regime the full-size 8K frame exposed: a camera hundreds of metres from the origin, a small near distance and a
    viewport thousands of pixels wide make the near corners of a tile sub-millimetre apart at large coordinates, so the planes
    the reference builds from them are tilted by up to degrees -- numerically, per tile.  The per-tile test reproduces that bit
    for bit; the two-level builder's conservative macro-cell pre-filter has to allow for it (it measures the tilt per macro cell).
    Wide, short viewports keep the oracle's O(tiles x lights) loop at a few seconds."""
    rng = np.random.default_rng(17000 + seed)
    w = int(rng.choice([3840, 7680, 5120, 2560]))
    h = int(rng.choice([64, 96, 160]))
    dist = float(rng.choice([150.0, 400.0, 1200.0]))
    zn = float(rng.choice([0.05, 0.1, 0.5]))
    zf = float(rng.choice([2000.0, 5000.0]))
    n = int(rng.integers(200, 700))
    ext = dist * 0.4
    lights = scenes.make_lights(n - n // 4, n // 4, (-ext, -0.1 * ext, -ext), (ext, 0.1 * ext, ext), seed=seed, range_lo=0.01 * dist, range_hi=0.06 * dist,
                                jolt_bounds=bool(seed % 2))
    ang = rng.uniform(0, 2 * np.pi)
    eye = (dist * np.sin(ang) + float(rng.uniform(-1000, 1000)) * (seed % 3 == 0), 0.25 * dist, -dist * np.cos(ang))
    off = (eye[0] - dist * np.sin(ang), 0.0, 0.0)
    lights["position_range"][:, 0] += np.float32(off[0]); lights["cull_sphere"][:, 0] += np.float32(off[0])
    lights["cull_aabb_min"][:, 0] += np.float32(off[0]); lights["cull_aabb_max"][:, 0] += np.float32(off[0])
    vp = scenes.camera_viewproj(tuple(float(v) for v in eye), (float(off[0]), 0.0, 0.0), (0.0, 1.0, 0.0), float(np.radians(rng.uniform(8, 25))), w / h, zn, zf)
    gpu.lights_upload(lights.view(np.uint8))
    gpu.light_cull(vp, w, h, 16, 128)
    c, i = gpu.light_lists_download()
    oc, oi = port.light_cull(lights, vp, w, h, 16, 128)
    assert int(oc.sum()) > 0
    keep = np.arange(128)[None, :] < np.minimum(oc, 128)[:, None]
    assert np.array_equal(c, oc), f"seed {seed}: counts differ in {int(np.count_nonzero(c != oc))} of {c.size} tiles ({w}x{h}, camera at {dist} m)"
    assert np.array_equal(i[keep], oi[keep]), f"seed {seed}: lists differ"


@pytest.mark.parametrize("seed", list(range(24)) + list(range(100, 116)))
def test_fuzz_specialised_tile_instantiations_parity(gpu, port, seed):
    """The random scenes again WITHOUT AOVs, so that the frames run the specialised instantiations of the tile kernel (Forward+ over
    point / spot lights or with area lights, no local lights with or without the sun shadow map, both programs) instead of the general
    one: depth <= 1 ULP, light lists and statistics exact, LDR <= 1 LSB, HDR >= 60 dB against the oracle."""
    lights = seed >= 100
    sd = scenes.scene_fuzz(seed, lights=lights, zero_normals=False)
    shadow = bool(sd.fp.shadow_enable) and not lights
    kw = {"forward_plus": True, "fused": bool(seed & 1)} if lights else {"shadow": shadow}
    g = harness.gpu_forward(gpu, sd, aov=False, **kw)
    mode = gpu.last_tile_kernel()
    kw.pop("fused", None)
    c = harness.cpu_forward(port, sd, aov=False, **kw)
    harness.assert_frame_parity(g, c, name=f"{sd.name} (instantiation {mode})")
    print(f"{sd.name}: tile kernel instantiation {mode}")
    FAST_SEEN.add(mode)


FAST_SEEN = set()


def test_fuzz_specialised_instantiations_were_exercised():
    """Runs after the parametrised test above (same file, same process): the seeds must have reached the Forward+ and the no-lights
    instantiations of both programs, not only the general kernel."""
    assert {11, 12} <= FAST_SEEN and (FAST_SEEN & {21, 22}), FAST_SEEN
