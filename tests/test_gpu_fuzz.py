"""GPU differential fuzzing (`-m gpu`): the CUDA path through the C-ABI against the CPU oracle on the seeded random scenes of
scenes.scene_fuzz (the same generator tests/test_fuzz_cpu.py uses to pin the oracle against the compiled reference, bit for bit,
on the CPU).  Gates as everywhere: coverage masks, triangle ids, fragment counts, statistics and light lists bit-exact, depth and
shadow depth <= 1 ULP, LDR <= 1 LSB, HDR PSNR >= 60 dB."""
import numpy as np
import pytest

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
