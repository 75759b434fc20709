"""Differential fuzzing on the CPU: the restatement (oracle/oracle.cpp) against the reference's own headers compiled into
oracle/_ref/libshs_ref.so on seeded random scenes (leisure_software_renderer_b200/scenes.py: scene_fuzz) -- random target
sizes, cameras inside / grazing geometry, depth ranges, shading models, cull modes and windings, mirrored and non-uniform
instances, triangle soups with slivers and frustum-crossing triangles, odd texture sizes, both sky models, random PCF
parameters.  Same machine, same libm, no FMA: every plane must be BIT-equal."""
import numpy as np
import pytest

import harness
from leisure_software_renderer_b200 import scenes

SEEDS = list(range(300))


def _same_bits(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint8), np.ascontiguousarray(b).view(np.uint8))


@pytest.mark.parametrize("seed", SEEDS)
def test_fuzz_forward_bit_exact(port, reference, seed):
    sd = scenes.scene_fuzz(seed)
    shadow = bool(sd.fp.shadow_enable)
    depth = seed % 7 != 3                       # every 7th scene without a depth target: painter's order
    a = harness.cpu_forward(port, sd, aov=False, shadow=shadow, depth=depth)
    b = harness.cpu_forward(reference, sd, aov=False, shadow=shadow, depth=depth)
    for k in ("tri_input", "tri_after_clip", "tri_raster"):
        assert a.stats[k] == b.stats[k], (sd.name, k, a.stats[k], b.stats[k])
    if depth:
        assert _same_bits(a.depth, b.depth), f"{sd.name}: depth differs at {int(np.count_nonzero(a.depth != b.depth))} px"
    if shadow:
        assert _same_bits(a.lvp, b.lvp), f"{sd.name}: light camera differs"
        assert _same_bits(a.shadow, b.shadow), f"{sd.name}: shadow map differs"
    assert _same_bits(a.hdr, b.hdr), f"{sd.name}: HDR differs at {int(np.count_nonzero((a.hdr != b.hdr).any(axis=2)))} px"
    assert _same_bits(a.ldr, b.ldr), f"{sd.name}: LDR differs"


def test_fuzz_scenes_are_not_trivial(port):
    """The generator must actually exercise the path: most scenes draw something, some clip, some are culled."""
    drawn = clipped = 0
    for seed in SEEDS[:16]:
        sd = scenes.scene_fuzz(seed)
        f = harness.cpu_forward(port, sd, aov=False, shadow=False, tonemap=False)
        drawn += f.stats["tri_raster"] > 0
        clipped += f.stats["tri_after_clip"] != f.stats["tri_input"]
    assert drawn >= 10 and clipped >= 6, (drawn, clipped)
