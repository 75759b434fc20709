"""CPU suite for the depth-range / clustered bin builders.  The reference's headers for these
(lighting/jolt_light_culling.hpp) need JoltPhysics; they are pinned in tests/test_light_cull_pinned_cpu.py (compiled against a
JoltPhysics declaration shim).  Here: structural properties that follow from the reference's definition, the per-tile depth reduce
(defined by this repository: the reference has it only as GLSL), and the fixture for the GPU box."""
import os

import numpy as np

import harness
from leisure_software_renderer_b200 import capi, scenes

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
