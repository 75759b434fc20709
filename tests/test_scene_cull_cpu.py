"""Scene-level steps upstream of draw submission on the CPU: the restatement
(oracle/oracle_scene_cull.cpp) PINNED against the reference's own cull_vs_frustum (geometry/jolt_culling.hpp:279-306) and
collect_object_lights (lighting/light_runtime.hpp:592-616), compiled where they lie against the JoltPhysics declaration shim
(oracle/ref_lightcull_harness.cpp) -- classes, the ordered visible list, the counters, and per object the selected light slots and
squared distances bit for bit -- and the DEVICE functions of csrc/scene_cull_core.cuh compiled by g++ (tests/cpp/scene_cull_emul.cpp)
against the restatement."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import fuzz_cases
from leisure_software_renderer_b200 import capi
from oracle.bindings import LightCullReference, SceneCull

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "..", "leisure_software_renderer_b200", "csrc")


@pytest.fixture(scope="module")
def refs():
    if not LightCullReference.available() and not os.path.isdir("/root/reference"):
        pytest.skip("oracle/_ref/libshs_lightcull_ref.so not built and /root/reference absent")
    return SceneCull("port"), SceneCull("reference"), LightCullReference()


class Emul(SceneCull):
    def __init__(self):
        out, src, hdr = os.path.join(HERE, "cpp", "_build", "libscene_cull_emul.so"), os.path.join(HERE, "cpp", "scene_cull_emul.cpp"), os.path.join(CSRC, "scene_cull_core.cuh")
        if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
            os.makedirs(os.path.dirname(out), exist_ok=True)
            subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-I" + CSRC, src, "-o", out], check=True)
        self.kind, self.lib, self.prefix = "port", C.CDLL(out), "shsemu_"


def object_bounds(light_ref, aabbs):
    """(n, 10) bounds as SceneShape reports them for these AABBs (the sphere is the reference's derivation)."""
    return light_ref.bounds(aabbs)


@pytest.mark.parametrize("seed", list(range(50)))
def test_fuzz_object_culling_equals_the_reference(refs, seed):
    port, ref, lref = refs
    sc = fuzz_cases.scene_cull(seed)
    b = object_bounds(lref, sc["aabbs"])
    pc, pv, pn = port.cull_objects(b, sc["view_proj"])
    rc, rv, rn = ref.cull_objects(sc["aabbs"], sc["view_proj"])
    assert np.array_equal(pc, rc) and np.array_equal(pv, rv) and np.array_equal(pn, rn), (seed, pn, rn)
    ec, ev, en = Emul().cull_objects(b, sc["view_proj"])
    assert np.array_equal(ec, pc) and np.array_equal(ev, pv) and np.array_equal(en, pn)


@pytest.mark.parametrize("seed", list(range(50)))
def test_fuzz_object_light_selection_equals_the_reference(refs, seed):
    port, ref, _ = refs
    sc = fuzz_cases.scene_cull(seed)
    for mode in (0, 1, 2):
        p = port.collect_object_lights(sc["aabbs"], sc["visible"], sc["lights"], mode)
        r = ref.collect_object_lights(sc["aabbs"], sc["visible"], sc["lights"], mode)
        e = Emul().collect_object_lights(sc["aabbs"], sc["visible"], sc["lights"], mode)
        for k, what in enumerate(("counts", "indices", "dist2")):
            assert np.array_equal(p[k].view(np.uint32), r[k].view(np.uint32)), f"seed {seed} mode {mode}: {what} differ from the reference"
            assert np.array_equal(e[k].view(np.uint32), p[k].view(np.uint32)), f"seed {seed} mode {mode}: device functions: {what} differ"


@pytest.mark.parametrize("seed", list(range(50)))
def test_fuzz_scene_tile_depth_range_equals_the_reference(refs, seed):
    """build_tile_view_depth_range_from_scene (lighting/light_culling_runtime.hpp:188-264): restatement == compiled reference, and the
    device functions (projection per object, ordered-key min / max folded in REVERSE object order, closing pass) == restatement."""
    port, ref, _ = refs
    sc = fuzz_cases.scene_cull(seed)
    args = (sc["aabbs"], sc["visible_objects"], sc["view"], sc["view_proj"], sc["w"], sc["h"], sc["ts"], sc["zn"], sc["zf"])
    p, r, e = port.tile_depth_range_from_scene(*args), ref.tile_depth_range_from_scene(*args), Emul().tile_depth_range_from_scene(*args)
    for k in (0, 1):
        assert np.array_equal(p[k].view(np.uint32), r[k].view(np.uint32)), f"seed {seed}: {'min' if k == 0 else 'max'} differs from the reference in {int(np.count_nonzero(p[k] != r[k]))} tiles"
        assert np.array_equal(e[k].view(np.uint32), p[k].view(np.uint32)), f"seed {seed}: device functions differ"


def test_scene_tile_depth_ranges_are_not_trivial(refs):
    port = refs[0]
    tight = full = 0
    for seed in range(20):
        sc = fuzz_cases.scene_cull(seed)
        lo, hi = port.tile_depth_range_from_scene(sc["aabbs"], sc["visible_objects"], sc["view"], sc["view_proj"], sc["w"], sc["h"], sc["ts"], sc["zn"], sc["zf"])
        assert np.all(lo <= hi)
        tight += int(np.count_nonzero((lo > np.float32(sc["zn"])) | (hi < np.float32(sc["zf"]))))
        full += int(np.count_nonzero((lo == np.float32(sc["zn"])) & (hi == np.float32(sc["zf"]))))
    assert tight > 500 and full > 500, (tight, full)


def test_scene_cull_scenes_are_not_trivial(refs):
    port, _, lref = refs
    seen = np.zeros(3, np.int64)
    full = replaced = 0
    for seed in range(20):
        sc = fuzz_cases.scene_cull(seed)
        cls, _, _ = port.cull_objects(object_bounds(lref, sc["aabbs"]), sc["view_proj"])
        seen += np.bincount(cls, minlength=3)
        counts, idx, d2 = port.collect_object_lights(sc["aabbs"], sc["visible"], sc["lights"], 1)
        full += int((counts == 8).sum())
        none = port.collect_object_lights(sc["aabbs"], sc["visible"], sc["lights"], 0)
        replaced += int(np.count_nonzero((none[1] != idx).any(axis=1)))
    assert (seen > 50).all(), seen
    assert full > 100 and replaced > 100, (full, replaced)


def _bins_for(sc, lref, culling_mode, slices):
    """The light bins of build_light_bin_culling for this scene in the capped device layout, from the PINNED bin builders of the
    restatement (cap = n_lights, so nothing is cut), plus the reference-derived light records and AABBs."""
    import test_light_cull_pinned_cpu as lp
    from oracle.bindings import Oracle
    recs, light_aabbs = lp.with_reference_bounds(lref, sc["lights"])
    zn32 = np.float32(max(np.float32(sc["zn"]), np.float32(1e-4)))
    zf32 = np.float32(max(np.float32(sc["zf"]), zn32 + np.float32(1e-3)))
    ts = max(sc["ts"], 1)
    mode = {1: capi.LIGHT_CULL_TILED, 2: capi.LIGHT_CULL_TILED_VIEW_DEPTH, 3: capi.LIGHT_CULL_CLUSTERED}[culling_mode]
    desc = capi.LightCullDesc(sc["view_proj"], sc["w"], sc["h"], mode, ts, len(recs), depth_slices=max(slices, 1), z_near=float(zn32), z_far=float(zf32))
    return recs, light_aabbs, desc, float(zn32), float(zf32), Oracle("port")


@pytest.mark.parametrize("seed", list(range(40)))
def test_fuzz_bin_gather_and_selection_equals_the_reference(refs, seed):
    """build_light_bin_culling -> gather_light_scene_candidates_for_aabb -> collect_object_lights, per object, as the reference's
    software light-culling demo chains them (exp-plumbing/hello_light_types_culling_sw.cpp:968-996): candidate counts, selected slots
    and squared distances of the restatement (fed with the pinned restatement's bins) and of the device functions, against the
    compiled reference, bit for bit -- tiled, tiled + scene depth ranges, clustered."""
    port, ref, lref = refs
    sc = fuzz_cases.scene_cull(seed)
    slices = [1, 3, 16][seed % 3]
    objs = sc["aabbs"][:120]
    for culling_mode in (1, 2, 3):
        recs, light_aabbs, desc, zn, zf, oport = _bins_for(sc, lref, culling_mode, slices)
        lo = hi = None
        if culling_mode == 2:
            lo, hi = port.tile_depth_range_from_scene(sc["aabbs"], sc["visible_objects"], sc["view"], sc["view_proj"], sc["w"], sc["h"], desc.tile_size, zn, zf)
        bc, bi = oport.light_cull_ex(recs, desc, lo, hi)
        tiles_x, tiles_y = (sc["w"] + desc.tile_size - 1) // desc.tile_size, (sc["h"] + desc.tile_size - 1) // desc.tile_size
        bins = (tiles_x, tiles_y, desc.depth_slices if culling_mode == 3 else 1)
        for cull_mode in (1, 2):
            r = ref.reference_select_object_lights(objs, sc["view"], sc["view_proj"], sc["w"], sc["h"], culling_mode, desc.tile_size, slices, sc["zn"], sc["zf"], lo, hi,
                                                   light_aabbs, recs, cull_mode)
            for who, impl in (("restatement", port), ("device functions", Emul())):
                p = impl.select_object_lights_from_bins(objs, sc["view"], sc["view_proj"], bins, culling_mode == 3, zn, zf, bc, bi, recs, cull_mode)
                for k, what in enumerate(("counts", "indices", "dist2", "candidates")):
                    assert np.array_equal(p[k].view(np.uint32), r[k].view(np.uint32)), f"seed {seed} bins {culling_mode} cull {cull_mode}: {who}: {what} differ"


def test_bin_gather_is_not_trivial(refs):
    port, _, lref = refs
    fewer = fallback = 0
    for seed in range(12):
        sc = fuzz_cases.scene_cull(seed)
        recs, light_aabbs, desc, zn, zf, oport = _bins_for(sc, lref, 3, 16)
        bc, bi = oport.light_cull_ex(recs, desc)
        tiles_x, tiles_y = (sc["w"] + desc.tile_size - 1) // desc.tile_size, (sc["h"] + desc.tile_size - 1) // desc.tile_size
        p = port.select_object_lights_from_bins(sc["aabbs"], sc["view"], sc["view_proj"], (tiles_x, tiles_y, 16), True, zn, zf, bc, bi, recs, 1)
        fewer += int(np.count_nonzero(p[3] < len(recs)))
        fallback += int(np.count_nonzero(p[3] == len(recs)))
    assert fewer > 300 and fallback > 20, (fewer, fallback)
