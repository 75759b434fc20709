"""GPU parity (`-m gpu`) of the flat-shaded mesh draws behind the C-ABI (shsb_flat_draw_blinn_phong, shsb_flat_draw_multi_light;
csrc/flat_draw.cu; SURVEY.md section 8f row 1: the consumer of the per-object light selections) against the CPU oracle
(oracle/oracle_flat_draw.cpp, pinned byte for byte against the reference's own text by tests/test_flat_draw_cpu.py):
depth buffer bit for bit, canvas within 1 LSB per channel with the number of differing channels reported (the device evaluates
std::pow / std::cos in double and narrows; expected and so far observed: 0)."""
import ctypes as C

import numpy as np
import pytest

import fuzz_cases
from leisure_software_renderer_b200 import capi
from oracle.bindings import FlatDraw, LightCullReference, SceneCull

pytestmark = pytest.mark.gpu
BATCH_KEYS = ("draw_mesh", "models", "base", "sel_counts", "sel_idx")


class Targets:
    """Meshes of a flat-draw scene uploaded one by one (positions from the mesh's base vertex on) + a canvas and a depth target."""

    def __init__(self, gpu, sc):
        self.gpu, self.sc = gpu, sc
        self.meshes = []
        for first, count, base_v in np.asarray(sc["mesh_table"]).reshape(-1, 3):
            self.meshes.append(gpu.mesh_upload(sc["vertices"][int(base_v):], indices=sc["indices"][int(first):int(first) + int(count)]))
        self.canvas = gpu.rt_create(capi.RT_COLOR_LDR, sc["W"], sc["H"])
        self.depth = gpu.rt_create(capi.RT_SHADOW, sc["W"], sc["H"])
        self.reset()

    def reset(self):
        self.gpu.rt_upload(self.canvas, capi.PLANE_COLOR, self.sc["canvas"])
        self.gpu.rt_upload(self.depth, capi.PLANE_DEPTH, self.sc["depth"])

    def draws(self, lo=0, hi=None):
        sc = self.sc
        hi = len(sc["draw_mesh"]) if hi is None else hi
        return [{"mesh": self.meshes[int(sc["draw_mesh"][i])], "model": sc["models"][i], "base_color": sc["base"][i],
                 "selection": sc["sel_idx"][i], "selection_count": int(sc["sel_counts"][i])} for i in range(lo, hi)]

    def run(self, mode, lo=0, hi=None):
        sc = self.sc
        if mode == 0:
            self.gpu.flat_draw_blinn_phong(self.draws(lo, hi), sc["view_proj"], sc["camera"], sc["light_dir"], self.canvas, self.depth)
        else:
            self.gpu.flat_draw_multi_light(self.draws(lo, hi), sc["view_proj"], sc["camera"], sc["lights"], self.canvas, self.depth)

    def read(self):
        return self.gpu.rt_download(self.canvas, capi.PLANE_COLOR).reshape(self.sc["H"], self.sc["W"], 4), self.gpu.rt_download(self.depth, capi.PLANE_DEPTH).reshape(self.sc["H"], self.sc["W"])

    def close(self):
        self.gpu.rt_destroy(self.canvas)
        self.gpu.rt_destroy(self.depth)
        for m in self.meshes:
            self.gpu.lib.shsb_mesh_destroy(self.gpu.h, m)


def compare(got, want, what):
    (gc, gd), (wc, wd) = got, want
    bad = int(np.count_nonzero(gd.view(np.uint32) != wd.view(np.uint32)))
    assert bad == 0, f"{what}: depth differs at {bad} texels"
    diff = np.abs(gc.astype(np.int16) - wc.astype(np.int16))
    n = int(np.count_nonzero(diff))
    assert diff.max() <= 1, f"{what}: canvas differs by {int(diff.max())} ({n} channels)"
    return n


@pytest.mark.parametrize("seed", list(range(40)))
def test_fuzz_flat_draws_equal_the_oracle(gpu, seed):
    sc = fuzz_cases.flat_draw_scene(seed, dangling=seed % 2 == 1)
    t = Targets(gpu, sc)
    try:
        off = 0
        for mode in (0, 1):
            t.reset()
            t.run(mode)
            off += compare(t.read(), FlatDraw("port").run(sc, mode), f"seed {seed} mode {mode}")
        print(f"seed {seed}: {off} colour channels off by 1 LSB")
        assert off <= 4
    finally:
        t.close()


@pytest.mark.parametrize("seed", [2, 11, 17])
def test_batches_compose_like_the_serial_loop(gpu, seed):
    """Two calls over the halves of a batch (the second sees the first's depth buffer) == one call == the oracle's serial loop."""
    sc = fuzz_cases.flat_draw_scene(seed)
    n = len(sc["draw_mesh"])
    t = Targets(gpu, sc)
    try:
        t.run(1, 0, n // 2)
        t.run(1, n // 2, n)
        compare(t.read(), FlatDraw("port").run(sc, 1), f"seed {seed} split batch")
    finally:
        t.close()


def test_demo_sized_frame_and_work_list_growth(gpu):
    """The demo's canvas (1200 x 900, hello_light_types_culling_sw.cpp:43-46) with 600 objects: the floor alone is ~270 work items, the
    first call has to grow its work list; compared in full with the oracle."""
    sc = fuzz_cases.flat_draw_scene(5, size=(1200, 900))
    rng = np.random.default_rng(77)
    reps = 600 // len(sc["draw_mesh"]) + 1
    big = dict(sc)
    for k in BATCH_KEYS:
        big[k] = np.concatenate([sc[k]] * reps)[:600].copy()
    big["models"][:, 12:15] += rng.uniform(-6, 6, (600, 3)).astype(np.float32)
    t = Targets(gpu, big)
    try:
        for mode in (0, 1):
            t.reset()
            t.run(mode)
            compare(t.read(), FlatDraw("port").run(big, mode), f"demo-sized frame mode {mode}")
    finally:
        t.close()


def test_selection_chain_feeds_the_draw(gpu):
    """The demo's per-frame chain (hello_light_types_culling_sw.cpp:968-1013) on the device end to end: collect_object_lights per object
    -> draw_mesh_multi_light_transformed with that selection, against the same chain through the oracle."""
    sc = fuzz_cases.flat_draw_scene(9, size=(320, 180))
    n = len(sc["draw_mesh"])
    li = sc["lights"]
    # CullingLightGPU-like records for the selection step: only the world bounds matter there (sphere at the light with its range)
    cs = fuzz_cases.scene_cull(9)
    rec = np.resize(cs["lights"], (len(li),) + cs["lights"].shape[1:]) if len(li) else cs["lights"][:0]
    aabbs = np.zeros((n, 6), np.float32)
    for i in range(n):
        M = sc["models"][i].reshape(4, 4).T
        first, count, base_v = sc["mesh_table"][int(sc["draw_mesh"][i])]
        idx = sc["indices"][int(first):int(first) + int(count)]
        pts = (sc["vertices"][int(base_v) + idx] @ M[:3, :3].T) + M[:3, 3]
        aabbs[i, :3], aabbs[i, 3:] = pts.min(axis=0), pts.max(axis=0)
    visible = np.arange(len(rec), dtype=np.uint32)
    counts, idx8, _ = gpu.collect_object_lights(aabbs, visible, rec, 1)
    pc, pi, _ = SceneCull("port").collect_object_lights(aabbs, visible, rec, 1)
    assert np.array_equal(counts, pc) and np.array_equal(idx8, pi)
    chain = dict(sc)
    chain["sel_counts"], chain["sel_idx"] = counts, idx8
    t = Targets(gpu, chain)
    try:
        t.run(1)
        compare(t.read(), FlatDraw("port").run(chain, 1), "selection chain")
    finally:
        t.close()


def test_flat_draw_argument_checks(gpu):
    sc = fuzz_cases.flat_draw_scene(1, size=(32, 24))
    t = Targets(gpu, sc)
    other = gpu.rt_create(capi.RT_SHADOW, 16, 16)
    try:
        lib, f = gpu.lib, capi.fptr(np.zeros(16, np.float32))
        d = gpu._flat_draws(t.draws(0, 1))
        assert lib.shsb_flat_draw_blinn_phong(gpu.h, d, 1, f, f, f, t.canvas, other) == 7            # SHSB_E_SIZE_MISMATCH
        assert lib.shsb_flat_draw_blinn_phong(gpu.h, d, 1, f, f, f, t.depth, t.depth) == 2           # canvas is not an RT_ColorLDR
        assert lib.shsb_flat_draw_blinn_phong(gpu.h, d, 1, None, f, f, t.canvas, t.depth) == 1
        assert lib.shsb_flat_draw_multi_light(gpu.h, d, 1, f, f, None, 3, t.canvas, t.depth) == 1    # lights null with a count
        d[0].selection_count = 9
        assert lib.shsb_flat_draw_multi_light(gpu.h, d, 1, f, f, None, 0, t.canvas, t.depth) == 1
        d[0].selection_count, d[0].mesh = 0, 9999
        assert lib.shsb_flat_draw_multi_light(gpu.h, d, 1, f, f, None, 0, t.canvas, t.depth) == 2
        assert lib.shsb_flat_draw_multi_light(gpu.h, None, 0, f, f, None, 0, t.canvas, t.depth) == 0  # empty batch
        before = t.read()
        t.run(0, 0, 0)
        after = t.read()
        assert np.array_equal(before[0], after[0]) and np.array_equal(before[1], after[1])
    finally:
        gpu.rt_destroy(other)
        t.close()


def test_flat_draw_drop_in_parity_on_gpu():
    """shs::b200::FlatDrawBatch over the reference's own types (DebugMesh, RT_ColorLDR, LightInstance with its four light models,
    LightSelection) vs the reference's per-object draws (tests/cpp/flat_draw_drop_in_test.cpp): depth bit for bit, canvas <= 1 LSB."""
    import os
    import subprocess
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cpp", "_build", "flat_draw_drop_in_test")
    if not os.path.exists(path):
        pytest.skip("tests/cpp/_build/flat_draw_drop_in_test was not built (needs /root/reference at build time)")
    r = subprocess.run([path], capture_output=True, text=True, timeout=300)
    print(r.stdout, r.stderr)
    assert r.returncode == 0, r.stdout + r.stderr
