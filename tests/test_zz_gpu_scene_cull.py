"""GPU parity (`-m gpu`) of the scene-level culling steps behind the C-ABI (shsb_cull_objects_frustum, shsb_collect_object_lights;
csrc/scene_cull.cu) against the CPU oracle (oracle/oracle_scene_cull.cpp, pinned bit for bit against the
reference's own cull_vs_frustum / collect_object_lights by tests/test_scene_cull_cpu.py): classes, the ordered visible list, the
counters, and per object the selected light slots and squared distances, all bit-exact.
(Sorts last on purpose: written after the round's GPU budget was spent, checked so far through the CPU emulation of its device
functions only.)"""
import numpy as np
import pytest

import fuzz_cases
from oracle.bindings import LightCullReference, SceneCull

pytestmark = pytest.mark.gpu


def bounds_of(aabbs):
    """(n, 10) sphere + AABB as SceneShape reports them: centre and half-extent length of the AABB."""
    if LightCullReference.available():
        return LightCullReference().bounds(aabbs)
    a = np.ascontiguousarray(aabbs, np.float32)
    half = np.float32(0.5) * (a[:, 3:] - a[:, :3])
    r = np.sqrt((half[:, 0] * half[:, 0] + half[:, 1] * half[:, 1]) + half[:, 2] * half[:, 2]).astype(np.float32)
    return np.concatenate([np.float32(0.5) * (a[:, :3] + a[:, 3:]), r[:, None], a], axis=1).astype(np.float32)


@pytest.mark.parametrize("seed", list(range(30)))
def test_fuzz_object_culling_parity(gpu, seed):
    sc = fuzz_cases.scene_cull(seed)
    b = bounds_of(sc["aabbs"])
    g = gpu.cull_objects_frustum(b, sc["view_proj"])
    c = SceneCull("port").cull_objects(b, sc["view_proj"])
    assert np.array_equal(g[0], c[0]) and np.array_equal(g[1], c[1]) and np.array_equal(g[2], c[2]), (seed, g[2], c[2])


@pytest.mark.parametrize("seed", list(range(30)))
def test_fuzz_object_light_selection_parity(gpu, seed):
    sc = fuzz_cases.scene_cull(seed)
    for mode in (0, 1, 2):
        g = gpu.collect_object_lights(sc["aabbs"], sc["visible"], sc["lights"], mode)
        c = SceneCull("port").collect_object_lights(sc["aabbs"], sc["visible"], sc["lights"], mode)
        for k, what in enumerate(("counts", "indices", "dist2")):
            assert np.array_equal(g[k].view(np.uint32), c[k].view(np.uint32)), f"seed {seed} mode {mode}: {what}"


def test_scene_cull_large_and_empty(gpu):
    """More objects than one compaction round (1024) and the empty cases."""
    rng = np.random.default_rng(3)
    sc = fuzz_cases.scene_cull(4)
    c = rng.uniform(-40, 40, (5000, 3))
    half = np.abs(rng.normal(0, 2.0, (5000, 3)))
    aabbs = np.concatenate([c - half, c + half], axis=1).astype(np.float32)
    b = bounds_of(aabbs)
    g, p = gpu.cull_objects_frustum(b, sc["view_proj"]), SceneCull("port").cull_objects(b, sc["view_proj"])
    assert np.array_equal(g[0], p[0]) and np.array_equal(g[1], p[1]) and np.array_equal(g[2], p[2]) and 0 < int(p[2][4]) < 5000
    g, p = gpu.collect_object_lights(aabbs, sc["visible"], sc["lights"], 1), SceneCull("port").collect_object_lights(aabbs, sc["visible"], sc["lights"], 1)
    assert all(np.array_equal(x.view(np.uint32), y.view(np.uint32)) for x, y in zip(g, p))
    e = gpu.cull_objects_frustum(np.zeros((0, 10), np.float32), sc["view_proj"])
    assert e[0].size == 0 and e[1].size == 0 and not e[2].any()
    e = gpu.collect_object_lights(aabbs[:3], np.zeros(0, np.uint32), sc["lights"], 2)
    assert not e[0].any() and not e[1].any()
    with pytest.raises(Exception, match="status 1"):
        gpu.collect_object_lights(aabbs[:3], sc["visible"], sc["lights"], 7)


@pytest.mark.parametrize("seed", list(range(30)))
def test_fuzz_scene_tile_depth_range_parity(gpu, seed):
    """build_tile_view_depth_range_from_scene: the device's atomic min / max over ordered keys == the reference's serial loop."""
    sc = fuzz_cases.scene_cull(seed)
    args = (sc["aabbs"], sc["visible_objects"], sc["view"], sc["view_proj"], sc["w"], sc["h"], sc["ts"], sc["zn"], sc["zf"])
    g, c = gpu.tile_depth_range_from_scene(*args), SceneCull("port").tile_depth_range_from_scene(*args)
    assert np.array_equal(g[0].view(np.uint32), c[0].view(np.uint32)) and np.array_equal(g[1].view(np.uint32), c[1].view(np.uint32)), seed


def test_scene_depth_ranges_feed_the_view_depth_bin_builder(gpu):
    """The ranges left on the device are what shsb_light_cull_ex consumes with NULL range pointers."""
    from leisure_software_renderer_b200 import capi
    from oracle.bindings import Oracle
    sc = fuzz_cases.scene_cull(2)
    args = (sc["aabbs"], sc["visible_objects"], sc["view"], sc["view_proj"], sc["w"], sc["h"], sc["ts"], sc["zn"], sc["zf"])
    lo, hi = gpu.tile_depth_range_from_scene(*args)
    desc = capi.LightCullDesc(sc["view_proj"], sc["w"], sc["h"], capi.LIGHT_CULL_TILED_VIEW_DEPTH, sc["ts"], 64, z_near=sc["zn"], z_far=sc["zf"])
    gpu.lights_upload(sc["lights"])
    gc, gi = gpu.light_cull_ex(desc)
    pc, pi = Oracle("port").light_cull_ex(sc["lights"], desc, lo, hi)
    assert np.array_equal(gc, pc)
    keep = np.arange(64)[None, :] < np.minimum(pc, 64)[:, None]
    assert np.array_equal(gi[keep], pi[keep])


@pytest.mark.parametrize("seed", list(range(12)))
def test_fuzz_bin_gather_and_selection_parity(gpu, seed):
    """shsb_select_object_lights_from_bins over the bins the context built (tiled and clustered) == the restatement fed with the
    restatement's bins (which tests/test_scene_cull_cpu.py pins against the reference's build_light_bin_culling -> gather -> collect)."""
    from leisure_software_renderer_b200 import capi
    from oracle.bindings import Oracle
    sc = fuzz_cases.scene_cull(seed)
    lights = sc["lights"]
    zn = float(np.float32(max(np.float32(sc["zn"]), np.float32(1e-4))))
    zf = float(np.float32(max(np.float32(sc["zf"]), np.float32(zn) + np.float32(1e-3))))
    gpu.lights_upload(lights)
    tiles_x, tiles_y = (sc["w"] + sc["ts"] - 1) // sc["ts"], (sc["h"] + sc["ts"] - 1) // sc["ts"]
    for clustered, slices in ((False, 1), (True, [1, 3, 16][seed % 3])):
        mode = capi.LIGHT_CULL_CLUSTERED if clustered else capi.LIGHT_CULL_TILED
        desc = capi.LightCullDesc(sc["view_proj"], sc["w"], sc["h"], mode, sc["ts"], max(1, len(lights)), depth_slices=slices, z_near=zn, z_far=zf)
        gpu.light_cull_ex(desc)
        bc, bi = Oracle("port").light_cull_ex(lights, desc)
        for cull_mode in (0, 1, 2):
            g = gpu.select_object_lights_from_bins(sc["aabbs"], sc["view"], sc["view_proj"], clustered, sc["zn"], sc["zf"], lights, cull_mode)
            c = SceneCull("port").select_object_lights_from_bins(sc["aabbs"], sc["view"], sc["view_proj"], (tiles_x, tiles_y, slices), clustered, zn, zf, bc, bi, lights, cull_mode)
            for k, what in enumerate(("counts", "indices", "dist2", "candidates")):
                assert np.array_equal(g[k].view(np.uint32), c[k].view(np.uint32)), f"seed {seed} clustered {clustered} cull {cull_mode}: {what}"


def test_scene_cull_drop_in_parity_on_gpu():
    """shs::b200::cull_vs_frustum / collect_object_lights / build_tile_view_depth_range_from_scene over the reference's own types vs the
    reference functions (tests/cpp/scene_cull_drop_in_test.cpp): every result equal."""
    import os
    import subprocess
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cpp", "_build", "scene_cull_drop_in_test")
    if not os.path.exists(path):
        pytest.skip("tests/cpp/_build/scene_cull_drop_in_test was not built (needs /root/reference at build time)")
    r = subprocess.run([path], capture_output=True, text=True, timeout=300)
    print(r.stdout, r.stderr)
    assert r.returncode == 0, r.stdout + r.stderr


@pytest.mark.parametrize("seed", list(range(40)))
def test_fuzz_software_occlusion_equals_the_oracle(gpu, seed):
    """shsb_software_occlusion (one persistent CTA walking the reference's front-to-back order, atomic-minimum occluder raster) against
    the pinned restatement: occluded flags, the ordered visible list, CullingStats counters and the occlusion depth buffer bit for bit,
    with occlusion on and off."""
    from oracle import bindings
    sc = fuzz_cases.occlusion_scene(seed)
    port = bindings.SoftwareOcclusion("port")
    for enable in (True, False):
        want = port.run(sc, enable)
        got = gpu.software_occlusion(sc["aabbs"], sc["visible"], sc["object_mesh"], sc["models"], sc["mesh_table"], sc["vertices"], sc["indices"], sc["view"], sc["view_proj"],
                                     sc["occ_w"], sc["occ_h"], sc["eps"], enable)
        for g, w, name in zip(got, want, ("occluded flags", "visible list", "counts", "depth buffer")):
            if name == "depth buffer" and not enable:
                continue
            assert np.array_equal(np.asarray(g) + 0, np.asarray(w) + 0), f"seed {seed} enable {enable}: {name} differ"


def test_software_occlusion_argument_checks(gpu):
    import ctypes as C
    from leisure_software_renderer_b200 import capi
    lib = gpu.lib
    z = np.zeros(16, np.float32)
    c4 = np.zeros(4, np.uint32)
    f = capi.fptr(z)
    assert lib.shsb_software_occlusion(gpu.h, None, 0, None, 0, None, None, None, 0, None, 0, None, 0, f, f, 0, 10, C.c_float(1e-4), 1, None, None, capi.u32ptr(c4), None) == 1   # empty buffer
    assert lib.shsb_software_occlusion(gpu.h, None, 0, None, 0, None, None, None, 0, None, 0, None, 0, f, f, 16, 8, C.c_float(1e-4), 1, None, None, capi.u32ptr(c4), None) == 0   # empty scene
    assert list(c4) == [0, 0, 0, 0]
