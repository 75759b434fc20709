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
