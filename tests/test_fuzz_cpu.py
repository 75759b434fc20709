"""Differential fuzzing on the CPU: the restatement (oracle/oracle.cpp) against the reference's own headers compiled into
oracle/_ref/libshs_ref.so on seeded random scenes (leisure_software_renderer_b200/scenes.py: scene_fuzz) -- random target
sizes, cameras inside / grazing geometry, depth ranges, shading models, cull modes and windings, mirrored and non-uniform
instances, triangle soups with slivers and frustum-crossing triangles, odd texture sizes, both sky models, random PCF
parameters.  Same machine, same libm, no FMA: every plane must be BIT-equal."""
import numpy as np
import pytest

import fuzz_cases
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


@pytest.mark.parametrize("seed", list(range(60)))
def test_fuzz_motion_vectors_bit_exact(port, reference, seed):
    """Two consecutive random frames: Context::history of frame 1 feeds the motion vectors of frame 2 (rasterizer.hpp:295-307,
    388-411); random per-item motion, sometimes far beyond the 96-px clamp, sometimes a moved camera."""
    prev, cur = fuzz_cases.motion_pair(seed)
    pm = prev.models(port)
    a = harness.cpu_forward(port, cur, aov=False, motion=True, prev_models=pm)
    b = harness.cpu_forward(reference, cur, aov=False, motion=True, prev_models=pm)
    assert _same_bits(a.depth, b.depth) and _same_bits(a.hdr, b.hdr), f"{cur.name}: colour / depth differ"
    assert _same_bits(a.motion, b.motion), f"{cur.name}: motion differs at {int(np.count_nonzero(a.motion.view(np.uint32) != b.motion.view(np.uint32)))} components"


@pytest.mark.parametrize("seed", list(range(40)))
def test_fuzz_post_passes_bit_exact(port, reference, seed):
    """PassMotionBlur / PassLightShafts with random parameters (also outside their sane ranges) on random planes."""
    ldr, depth, motion, p, q, with_depth = fuzz_cases.post_inputs(seed)
    a = port.pass_motion_blur(p, ldr, motion, depth)
    b = reference.pass_motion_blur(p, ldr, motion, depth)
    assert np.array_equal(a, b), f"blur seed {seed}: {int(np.count_nonzero((a != b).any(axis=2)))} pixels differ"
    a = port.pass_light_shafts(q, ldr, depth if with_depth else None)
    b = reference.pass_light_shafts(q, ldr, depth if with_depth else None)
    assert np.array_equal(a, b), f"shafts seed {seed}: {int(np.count_nonzero((a != b).any(axis=2)))} pixels differ"


@pytest.mark.parametrize("seed", list(range(24)))
def test_fuzz_light_bins_properties(port, seed):
    """The bin builders have no compilable reference (jolt_light_culling.hpp needs JoltPhysics: parity unpinned), so the fuzzed
    restatement is held to the structure the reference's code implies: the plain builder == mode 0 of the _ex entry; strictly
    ascending lists; counts are uncapped and a capped call keeps exactly the first max_per_bin entries; a bin never lists a light
    the camera frustum rejects.  (Set relations between the modes -- a depth range only removes lights, a tile's clusters cover its
    plain list -- are properties of the geometry, not of the float arithmetic: a sub-cell's side planes are rebuilt from its own
    corners and differ from the tile's by more than the 1e-5 tolerance, so borderline lights flip.  They are checked on the
    well-conditioned scene of tests/test_light_bins_cpu.py, not under fuzz.)"""
    lights, descs = fuzz_cases.light_bins(seed)
    d0 = descs[0][1]
    pc, pi = port.light_cull(lights, np.array(list(d0.view_proj), np.float32), d0.viewport_w, d0.viewport_h, d0.tile_size, 1024)
    out = {}
    for name, d, lo, hi in descs:
        big = type(d).from_buffer_copy(d)
        big.max_per_bin = 1024                                   # uncapped view of the lists for the set relations
        out[name] = port.light_cull_ex(lights, big, lo, hi)
        c, i = out[name]
        for b in np.nonzero(c > 1)[0][:200]:
            row = i[b, : min(int(c[b]), 1024)]
            assert np.all(np.diff(row.astype(np.int64)) > 0), f"{name}: list of bin {b} is not strictly ascending"
        cc, ci = port.light_cull_ex(lights, d, lo, hi)          # the capped call keeps the uncapped counts and the first entries
        assert np.array_equal(cc, c)
        keep = np.arange(d.max_per_bin)[None, :] < np.minimum(c, d.max_per_bin)[:, None]
        assert np.array_equal(ci[keep], i[:, : d.max_per_bin][keep])
    assert np.array_equal(out["tiled"][0], pc) and np.array_equal(out["tiled"][1], pi)
    n_lights = len(np.frombuffer(lights.tobytes(), np.uint8)) // 160
    for name in out:
        c, i = out[name]
        assert int(c.max()) <= n_lights
        used = i[np.arange(1024)[None, :] < np.minimum(c, 1024)[:, None]]
        assert used.size == 0 or int(used.max()) < n_lights, f"{name}: index beyond the light set"


@pytest.mark.parametrize("seed", list(range(60)))
def test_fuzz_triangle_ids_through_reference_callbacks(port, reference, seed):
    """Triangle ids and per-pixel fragment counts are bit-exact gates of the north_star.  The reference has no such outputs: its
    side obtains them through its OWN shader-callback API (a counting VS wrapper + a recording FS wrapper around the builtin
    program, oracle/ref_harness.cpp), drawing the fuzz scene item by item with rasterize_mesh -- including the NON-indexed soups
    that PassPBRForward would skip as MeshData::empty()."""
    from oracle.bindings import HostAssets
    from leisure_software_renderer_b200 import capi
    from test_oracle_vs_reference import _uniforms
    sd = scenes.scene_fuzz(seed)
    depth = seed % 5 != 2
    A = HostAssets(sd.meshes, sd.textures)
    shader = capi.SHADER_BLINN_PHONG if sd.fp.shading_model == capi.SHADING_BLINN else capi.SHADER_PBR_MR
    outs = []
    for o in (port, reference):
        hdr = np.zeros((sd.h, sd.w, 4), np.float32)
        dep = np.ones((sd.h, sd.w), np.float32) if depth else None
        tri = np.full((sd.h, sd.w), capi.TRI_ID_NONE, np.uint32)
        cov = np.zeros((sd.h, sd.w), np.uint32)
        tgt = o.make_target(sd.w, sd.h, hdr, dep, tri_id=tri, coverage=cov, zn=sd.zn, zf=sd.zf)
        key = 0
        tot = {"tri_input": 0, "tri_after_clip": 0, "tri_raster": 0}
        for it in sd.items:
            if not it.get("visible", True):
                continue
            model = o.model_from_transform(it["pos"], it.get("rot", (0, 0, 0)), it.get("scl", (1, 1, 1)))
            mat = it.get("material") or {"base_color": (0.8, 0.5, 0.2), "metallic": 0.1, "roughness": 0.5}
            mesh = sd.meshes[it["mesh"] - 1]
            n_tris = (len(mesh["indices"]) if len(mesh["indices"]) else len(mesh["positions"])) // 3
            st = o.rasterize_mesh(A, it["mesh"], shader, _uniforms(sd, model, mat), tgt, key_base=key,
                                  cull_mode=sd.fp.cull_mode, front_face_ccw=bool(sd.fp.front_face_ccw))
            key += n_tris * 8
            for k in tot:
                tot[k] += getattr(st, k)
        outs.append((hdr, dep, tri, cov, tot))
    (h0, d0, t0, c0, s0), (h1, d1, t1, c1, s1) = outs
    assert s0 == s1, (sd.name, s0, s1)
    assert _same_bits(h0, h1), f"{sd.name}: HDR differs"
    if depth:
        assert _same_bits(d0, d1)
    assert np.array_equal(t0 >> 3, t1 >> 3), f"{sd.name}: {int(np.count_nonzero((t0 >> 3) != (t1 >> 3)))} pixels have another winning (item, triangle)"
    if not depth:
        assert np.array_equal(c0, c1), f"{sd.name}: fragment counts differ"
    else:
        assert np.all(c1 <= c0) and np.array_equal(c1 > 0, c0 > 0)
