"""The restatements (CPU; the CUDA path: tests/test_zz_gpu_golden_round1b.py) against tests/golden/golden_round1b.npz, written by
tests/golden/make_golden_round1b.py from the reference's OWN compiled code: legacy soft-shadow / PBR-IBL frames (rows L2 / L3), the four
light-list builders (A11, 8f row 2), object culling, per-object light selection and scene-based tile depth ranges (8f row 1)."""
import os
import sys

import numpy as np
import pytest

import fuzz_cases
import test_legacy2_cpu as t2
import test_legacy3_cpu as t3
from oracle.bindings import Legacy2Oracle, Legacy3Oracle, SceneCull

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
import make_golden_round1b as mg  # noqa: E402

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_round1b.npz"))


def same(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint8), np.ascontiguousarray(b).view(np.uint8))


def records_with(lights, bounds):
    r = np.array(lights, copy=True)
    r["cull_sphere"][:, :] = bounds[:, 0:4]
    r["cull_aabb_min"][:, :3] = bounds[:, 4:7]
    r["cull_aabb_max"][:, :3] = bounds[:, 7:10]
    return r


def test_restatements_match_the_reference_fixture(port):
    for seed in mg.L2_SEEDS:
        r = t2.render(Legacy2Oracle("port"), fuzz_cases.legacy2_scene(seed))
        assert all(same(x, G[f"l2_{seed}_{k}"]) for x, k in zip(r, ("shadow", "canvas", "z"))), seed
    for seed in mg.L3_SEEDS:
        r = t3.render(Legacy3Oracle("port"), fuzz_cases.legacy3_scene(seed))
        assert all(same(x, G[f"l3_{seed}_{k}"]) for x, k in zip(r, ("shadow", "canvas", "z", "velocity"))), seed
    for seed in mg.LIGHT_SEEDS:
        lights, descs = fuzz_cases.light_bins(seed)
        recs = records_with(lights, G[f"lights_{seed}_bounds"])
        for name, desc, lo, hi in descs:
            c, i = port.light_cull_ex(recs, desc, lo, hi)
            keep = np.arange(i.shape[1])[None, :] < np.minimum(c, i.shape[1])[:, None]
            assert same(c, G[f"lights_{seed}_{name}_counts"]) and np.array_equal(i[keep], G[f"lights_{seed}_{name}_indices"][keep]), (seed, name)
    sp = SceneCull("port")
    for seed in mg.SCENE_SEEDS:
        sc = fuzz_cases.scene_cull(seed)
        cls, vis, cnt = sp.cull_objects(G[f"scene_{seed}_bounds"], sc["view_proj"])
        assert same(cls, G[f"scene_{seed}_classes"]) and same(vis, G[f"scene_{seed}_visible"]) and same(cnt, G[f"scene_{seed}_counts"])
        for mode in (0, 1, 2):
            r = sp.collect_object_lights(sc["aabbs"], sc["visible"], sc["lights"], mode)
            assert all(same(x, G[f"scene_{seed}_sel{mode}_{k}"]) for x, k in zip(r, ("counts", "indices", "dist2"))), (seed, mode)
        lo, hi = sp.tile_depth_range_from_scene(sc["aabbs"], sc["visible_objects"], sc["view"], sc["view_proj"], sc["w"], sc["h"], sc["ts"], sc["zn"], sc["zf"])
        assert same(lo, G[f"scene_{seed}_range_min"]) and same(hi, G[f"scene_{seed}_range_max"])


def test_fixture_is_not_trivial():
    assert int((G["l2_2_z"] < 3e38).sum()) > 500 and int(np.count_nonzero(G["l3_4_velocity"])) > 100
    assert int(G["lights_12_tiled_counts"].sum()) > 1000 and int(G["lights_6_tiled_counts"].sum()) > 20 and int(G["scene_3_counts"][4]) > 5 and int(G["scene_3_sel1_counts"].sum()) > 50

