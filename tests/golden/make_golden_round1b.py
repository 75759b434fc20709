"""Generates tests/golden/golden_round1b.npz by running THE REFERENCE's OWN code (oracle/_ref/libshs_legacy2_ref.so, libshs_legacy3_ref.so,
libshs_lightcull_ref.so: hello_shadow_mapping_soft.cpp, hello_pbr.cpp, lighting/jolt_light_culling.hpp, geometry/jolt_culling.hpp,
lighting/light_runtime.hpp, lighting/light_culling_runtime.hpp compiled where they lie under /root/reference) on seeded scenes of
tests/fuzz_cases.py: the legacy soft-shadow and PBR / IBL frames (rows L2 / L3), the four light-list builders (row A11, 8f row 2),
object culling, per-object light selection and the scene-based tile depth ranges (8f row 1).  Run in the container that has
/root/reference; the committed fixture lets a box without the reference tree check the oracle and the CUDA path against reference
output."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))

import fuzz_cases  # noqa: E402
import test_legacy2_cpu as t2  # noqa: E402
import test_legacy3_cpu as t3  # noqa: E402
import test_light_cull_pinned_cpu as lp  # noqa: E402
from oracle.bindings import Legacy2Oracle, Legacy3Oracle, LightCullReference, SceneCull  # noqa: E402

L2_SEEDS, L3_SEEDS, LIGHT_SEEDS, SCENE_SEEDS = (2, 9), (4, 13), (1, 6, 12), (3, 8)


def main():
    out = {}
    for seed in L2_SEEDS:
        r = t2.render(Legacy2Oracle("reference"), fuzz_cases.legacy2_scene(seed))
        out[f"l2_{seed}_shadow"], out[f"l2_{seed}_canvas"], out[f"l2_{seed}_z"] = r
    for seed in L3_SEEDS:
        r = t3.render(Legacy3Oracle("reference"), fuzz_cases.legacy3_scene(seed))
        out[f"l3_{seed}_shadow"], out[f"l3_{seed}_canvas"], out[f"l3_{seed}_z"], out[f"l3_{seed}_velocity"] = r
    lref = LightCullReference()
    for seed in LIGHT_SEEDS:
        lights, descs = fuzz_cases.light_bins(seed)
        recs, aabbs = lp.with_reference_bounds(lref, lights)
        out[f"lights_{seed}_bounds"] = lref.bounds(aabbs)
        for name, desc, lo, hi in descs:
            c, i = lref.light_cull(aabbs, desc, lo, hi)
            out[f"lights_{seed}_{name}_counts"], out[f"lights_{seed}_{name}_indices"] = c, i
    sref = SceneCull("reference")
    for seed in SCENE_SEEDS:
        sc = fuzz_cases.scene_cull(seed)
        out[f"scene_{seed}_bounds"] = lref.bounds(sc["aabbs"])
        out[f"scene_{seed}_classes"], out[f"scene_{seed}_visible"], out[f"scene_{seed}_counts"] = sref.cull_objects(sc["aabbs"], sc["view_proj"])
        for mode in (0, 1, 2):
            c, i, d = sref.collect_object_lights(sc["aabbs"], sc["visible"], sc["lights"], mode)
            out[f"scene_{seed}_sel{mode}_counts"], out[f"scene_{seed}_sel{mode}_indices"], out[f"scene_{seed}_sel{mode}_dist2"] = c, i, d
        lo, hi = sref.tile_depth_range_from_scene(sc["aabbs"], sc["visible_objects"], sc["view"], sc["view_proj"], sc["w"], sc["h"], sc["ts"], sc["zn"], sc["zf"])
        out[f"scene_{seed}_range_min"], out[f"scene_{seed}_range_max"] = lo, hi
    np.savez_compressed(os.path.join(HERE, "golden_round1b.npz"), **out)
    print(len(out), "arrays,", os.path.getsize(os.path.join(HERE, "golden_round1b.npz")), "bytes")


if __name__ == "__main__":
    main()
