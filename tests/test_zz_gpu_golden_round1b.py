"""`-m gpu`: the CUDA paths written after the round's GPU budget was spent (legacy render-target demos, scene-level culling) directly
against the reference-generated fixture tests/golden/golden_round1b.npz (tests/golden/make_golden_round1b.py).  Sorts last like the
other test_zz files."""
import numpy as np
import pytest

import fuzz_cases
import test_golden_round1b as tg
from test_zz_gpu_legacy2 import check, gpu_render

pytestmark = pytest.mark.gpu
G, mg, same = tg.G, tg.mg, tg.same


def test_legacy_render_target_demos_match_the_reference_fixture(gpu):
    for seed in mg.L2_SEEDS:
        g = gpu_render(gpu, fuzz_cases.legacy2_scene(seed))
        check(g, [G[f"l2_{seed}_{k}"] for k in ("shadow", "canvas", "z")], f"L2 golden {seed}")
    for seed in mg.L3_SEEDS:
        g = gpu_render(gpu, fuzz_cases.legacy3_scene(seed), pbr=True)
        check(g, [G[f"l3_{seed}_{k}"] for k in ("shadow", "canvas", "z", "velocity")], f"L3 golden {seed}")


def test_scene_culling_matches_the_reference_fixture(gpu):
    for seed in mg.SCENE_SEEDS:
        sc = fuzz_cases.scene_cull(seed)
        cls, vis, cnt = gpu.cull_objects_frustum(G[f"scene_{seed}_bounds"], sc["view_proj"])
        assert same(cls, G[f"scene_{seed}_classes"]) and same(vis, G[f"scene_{seed}_visible"]) and same(cnt, G[f"scene_{seed}_counts"])
        for mode in (0, 1, 2):
            r = gpu.collect_object_lights(sc["aabbs"], sc["visible"], sc["lights"], mode)
            assert all(same(x, G[f"scene_{seed}_sel{mode}_{k}"]) for x, k in zip(r, ("counts", "indices", "dist2"))), (seed, mode)
        lo, hi = gpu.tile_depth_range_from_scene(sc["aabbs"], sc["visible_objects"], sc["view"], sc["view_proj"], sc["w"], sc["h"], sc["ts"], sc["zn"], sc["zf"])
        assert same(lo, G[f"scene_{seed}_range_min"]) and same(hi, G[f"scene_{seed}_range_max"])


def test_light_lists_match_the_reference_fixture(gpu):
    """The CUDA path of the four light-list builders (green on B200 against the restatement earlier in the round) directly against the
    reference-generated lists."""
    for seed in mg.LIGHT_SEEDS:
        lights, descs = fuzz_cases.light_bins(seed)
        recs = tg.records_with(lights, G[f"lights_{seed}_bounds"])
        gpu.lights_upload(recs)
        for name, desc, lo, hi in descs:
            c, i = gpu.light_cull_ex(desc, lo, hi)
            keep = np.arange(i.shape[1])[None, :] < np.minimum(c, i.shape[1])[:, None]
            assert same(c, G[f"lights_{seed}_{name}_counts"]) and np.array_equal(i[keep], G[f"lights_{seed}_{name}_indices"][keep]), (seed, name)
