"""CPU suite for the depth-range / clustered bin builders.  The reference's headers for these
(lighting/jolt_light_culling.hpp) need JoltPhysics; they are pinned in tests/test_light_cull_pinned_cpu.py (compiled against a
JoltPhysics declaration shim).  Here: structural properties that follow from the reference's definition, the per-tile depth reduce
(defined by this repository: the reference has it only as GLSL), and the fixture for the GPU box."""
import os

import numpy as np
import pytest

import harness
from leisure_software_renderer_b200 import capi, scenes
from oracle import bindings

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "golden_light_bins_port.npz")


def bins_scene():
    return scenes.scene_small(w=208, h=120, lights=96, seed=5)


def _lists(counts, indices):
    return [set(int(i) for i in indices[b, :min(int(c), indices.shape[1])]) for b, c in enumerate(counts)]


def descs(sd, mx=128):
    mk = lambda mode, **kw: capi.LightCullDesc(sd.viewproj, sd.w, sd.h, mode, 16, mx, z_near=sd.zn, z_far=sd.zf, **kw)
    return {"tiled": mk(capi.LIGHT_CULL_TILED), "depth01": mk(capi.LIGHT_CULL_TILED_DEPTH01), "view_depth": mk(capi.LIGHT_CULL_TILED_VIEW_DEPTH),
            "clustered": mk(capi.LIGHT_CULL_CLUSTERED, depth_slices=6)}


def test_bin_builders_structure(port):
    sd = bins_scene()
    f = harness.cpu_forward(port, sd, aov=False)
    lo, hi = port.tile_depth_range(f.depth, 16, sd.zn, sd.zf)
    d = descs(sd)
    base_c, base_i = port.light_cull(sd.lights, sd.viewproj, sd.w, sd.h, 16, 128)
    c0, i0 = port.light_cull_ex(sd.lights, d["tiled"])
    assert np.array_equal(c0, base_c) and np.array_equal(i0, base_i)                          # mode 0 is cull_lights_tiled
    # full-range depth arrays reproduce the plain tiled result: view depth [zn, zf] maps to NDC [-1, 1] only up to rounding,
    # depth01 [0, 1] maps exactly
    n = d["depth01"].tiles()
    c1, i1 = port.light_cull_ex(sd.lights, d["depth01"], np.zeros(n, np.float32), np.ones(n, np.float32))
    assert np.array_equal(c1, base_c) and np.array_equal(i1, base_i)
    # measured ranges: subsets of the plain lists, ascending, and strictly tighter somewhere
    c2, i2 = port.light_cull_ex(sd.lights, d["view_depth"], lo, hi)
    base_l, tight_l = _lists(base_c, base_i), _lists(c2, i2)
    covered = hi < np.float32(sd.zf)
    assert covered.any() and (~covered).any()
    for t in range(n):
        row = i2[t, :min(int(c2[t]), 128)]
        assert np.all(np.diff(row.astype(np.int64)) > 0)
        if lo[t] < hi[t]:                                                                   # a non-degenerate cell is a sub-frustum of the tile
            assert tight_l[t] <= base_l[t], t
    assert sum(len(s) for s in tight_l) < sum(len(s) for s in base_l)
    # empty tiles carry [zn, zf]
    assert np.all(lo[~covered] == np.float32(sd.zn)) and np.all(hi[~covered] == np.float32(sd.zf))
    # clustered: bins = slices * tiles; the union over the slices of a tile covers the tile's plain list
    c3, i3 = port.light_cull_ex(sd.lights, d["clustered"])
    assert c3.size == 6 * n
    cl = _lists(c3, i3)
    for t in range(n):
        union = set().union(*[cl[z * n + t] for z in range(6)])
        assert union >= base_l[t] or int(base_c[t]) > 128, t


def test_depth_range_definition(port):
    """shso_tile_depth_range against an independent numpy statement of its definition."""
    sd = bins_scene()
    f = harness.cpu_forward(port, sd, aov=False)
    lo, hi = port.tile_depth_range(f.depth, 16, sd.zn, sd.zf)
    tx, ty = (sd.w + 15) // 16, (sd.h + 15) // 16
    top = f.depth[::-1]                                                                      # tile rows count from the top
    zn, zf = np.float32(sd.zn), np.float32(sd.zf)
    for t in range(tx * ty):
        blk = top[(t // tx) * 16:(t // tx + 1) * 16, (t % tx) * 16:(t % tx + 1) * 16]
        sel = blk[blk < 1.0]
        if sel.size == 0:
            assert lo[t] == zn and hi[t] == zf
        else:
            vz = zn + sel * (zf - zn)
            assert lo[t] == vz.min() and hi[t] == vz.max()


def test_light_bins_golden(port):
    sd = bins_scene()
    g = np.load(GOLDEN)
    f = harness.cpu_forward(port, sd, aov=False)
    lo, hi = port.tile_depth_range(f.depth, 16, sd.zn, sd.zf)
    assert np.array_equal(lo, g["range_min"]) and np.array_equal(hi, g["range_max"])
    for name, d in descs(sd, mx=32).items():
        rng = (lo, hi) if name == "view_depth" else ((g["range01_min"], g["range01_max"]) if name == "depth01" else (None, None))
        c, i = port.light_cull_ex(sd.lights, d, *rng)
        assert np.array_equal(c, g[name + "_counts"]) and np.array_equal(i, g[name + "_indices"]), name


# ---------------------------------------------------------------------------------------------------------------------------------
# The per-tile depth reduce the reference HAS: shaders/vulkan/fp_stress_depth_reduce.comp, on projective depth in [0, 1].  Its text,
# lifted into C++ by oracle/extract_glsl_a9.py and compiled by oracle/ref_glsl_a9_harness.cpp, pins the restatement
# (shso_tile_depth_range_ndc01) bit for bit; the device path (shsb_tile_depth_range_ndc01) is held to the restatement by
# tests/test_gpu_light_bins.py.
def ndc01_depth_buffers(seed):
    """Depth planes in the hardware's zero-to-one encoding of an LH projection, plus what a caller may hand in: cleared texels (exactly 1), values above 1, negative values, whole tiles without geometry, ragged sizes."""
    rng = np.random.default_rng(31000 + seed)
    w, h = [(64, 48), (97, 53), (33, 17), (160, 90), (16, 16)][seed % 5]
    zn, zf = [(0.1, 1000.0), (0.05, 300.0), (1.0, 50.0), (0.0, 0.0), (5.0, 4.0)][(seed // 5) % 5]     # the last two exercise the shader's clamps of near / far
    n, f = max(zn, 0.001), max(zf, max(zn, 0.001) + 0.01)
    vz = rng.uniform(n, f, (h, w)) ** rng.choice([1.0, 0.3, 3.0])
    vz = np.clip(vz, n, f)
    d = (f / (f - n) - n * f / ((f - n) * vz)).astype(np.float32)                                   # zero-to-one depth of view depth vz
    d[rng.random((h, w)) < 0.35] = 1.0
    d[rng.random((h, w)) < 0.02] = np.float32(1.5)
    d[rng.random((h, w)) < 0.02] = np.float32(-0.25)
    d[rng.random((h, w)) < 0.02] = np.float32(0.0)
    d[rng.random((h, w)) < 0.02] = np.nextafter(np.float32(1.0), np.float32(0.0))
    ty, tx = int(rng.integers(0, max(1, h // 16))), int(rng.integers(0, max(1, w // 16)))
    d[h - 16 * (ty + 1): h - 16 * ty, 16 * tx: 16 * (tx + 1)] = 1.0                                  # a tile without geometry (rows are y-up)
    return d, float(zn), float(zf)


@pytest.mark.skipif(not bindings.glsl_depth_reduce_available(), reason="oracle/_ref/libshs_glsl_a9_ref.so not built and /root/reference absent")
@pytest.mark.parametrize("seed", list(range(50)))
def test_depth_reduce_restatement_equals_the_shader(port, seed):
    d, zn, zf = ndc01_depth_buffers(seed)
    for ts in (16, 8, 32):
        glo, ghi = bindings.glsl_tile_depth_range_ndc01(d, ts, zn, zf)
        plo, phi = port.tile_depth_range_ndc01(d, ts, zn, zf)
        assert np.array_equal(glo.view(np.uint32), plo.view(np.uint32)) and np.array_equal(ghi.view(np.uint32), phi.view(np.uint32)), (seed, ts)
    lo16, hi16 = port.tile_depth_range_ndc01(d, 16, zn, zf)
    assert np.count_nonzero((lo16 == 0) & (hi16 == 0)) >= 1 and (d.size <= 256 or np.count_nonzero(hi16 > lo16) >= 1)      # empty tiles and real ranges both occur


def test_depth_reduce_recovers_view_depths(port):
    """The encoding the shader inverts is the hardware's zero-to-one depth of the LH projection, d = f / (f - n) - n f / ((f - n) z):
    per tile the reduce gives back min / max of the view depths the plane was built from, to float accuracy.  (The software
    rasteriser never writes this encoding -- its fallback for zf <= zn is the interpolated CLIP z * 0.5 + 0.5, rasterizer.hpp:345-347 --
    so the plane comes from the caller: the Vulkan backend's depth attachment, or an upload.)"""
    rng = np.random.default_rng(5)
    n, f, w, h = 0.1, 200.0, 96, 64
    vz = rng.uniform(0.5, 150.0, (h, w))
    d = (f / (f - n) - n * f / ((f - n) * vz)).astype(np.float32)
    d[16:32, 32:48] = 1.0                                    # rows are y-up: tile row 2 from the top, tile column 2
    d[40:44, 0:90] = 1.0
    lo, hi = port.tile_depth_range_ndc01(d, 16, n, f)
    top, dtop = vz[::-1], d[::-1]
    empty = 0
    for t in range(24):
        rows, cols = slice((t // 6) * 16, (t // 6 + 1) * 16), slice((t % 6) * 16, (t % 6 + 1) * 16)
        sel = top[rows, cols][dtop[rows, cols] < 1]
        if sel.size == 0:
            assert lo[t] == 0 and hi[t] == 0
            empty += 1
        else:
            assert abs(lo[t] - sel.min()) <= 2e-2 * sel.min() and abs(hi[t] - sel.max()) <= 2e-2 * sel.max(), (t, lo[t], sel.min(), hi[t], sel.max())
    assert empty == 1
