"""CPU suite: pins the restatement (oracle/oracle.cpp) against the reference's own headers compiled into
oracle/_ref/libshs_ref.so.  Everything here must be BIT-exact: same machine, same libm, no FMA."""
import ctypes as C

import numpy as np
import pytest

import cases
import harness
from leisure_software_renderer_b200 import capi, scenes
from oracle.bindings import HostAssets


def _same_bits(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint8), np.ascontiguousarray(b).view(np.uint8))


@pytest.mark.parametrize("name", list(cases.forward_cases()))
def test_forward_pass_bit_exact(port, reference, name):
    make, kw = cases.forward_cases()[name]
    sd = make()
    a = harness.cpu_forward(port, sd, aov=False, **kw)
    b = harness.cpu_forward(reference, sd, aov=False, **kw)
    assert _same_bits(a.hdr, b.hdr), f"{name}: HDR differs"
    if a.depth is not None:
        assert _same_bits(a.depth, b.depth), f"{name}: depth differs"
    assert _same_bits(a.ldr, b.ldr), f"{name}: LDR differs"
    if a.shadow is not None:
        assert _same_bits(a.shadow, b.shadow) and _same_bits(a.lvp, b.lvp), f"{name}: shadow map / light camera differs"
    for k in ("tri_input", "tri_after_clip", "tri_raster"):
        assert a.stats[k] == b.stats[k], (name, k, a.stats[k], b.stats[k])


@pytest.mark.parametrize("name", list(cases.motion_cases()))
def test_motion_vectors_bit_exact(port, reference, name):
    """RT_ColorDepthMotion::motion (rasterizer.hpp:295-307, 388-411) with Context::history from a previous frame."""
    make, kw = cases.motion_cases()[name]
    prev, cur = make()
    pm = prev.models(port) if prev is not None else None
    a = harness.cpu_forward(port, cur, aov=False, motion=True, prev_models=pm, **kw)
    b = harness.cpu_forward(reference, cur, aov=False, motion=True, prev_models=pm, **kw)
    assert _same_bits(a.hdr, b.hdr) and _same_bits(a.depth, b.depth), f"{name}: colour / depth differ"
    assert _same_bits(a.motion, b.motion), f"{name}: motion differs at {int(np.count_nonzero(a.motion != b.motion))} components"
    covered = a.depth < 1.0
    assert np.all(a.motion[~covered] == 0.0), "the pass clears the motion plane where nothing is drawn"
    if prev is not None:
        assert np.count_nonzero(a.motion[covered]) > 0, "moving objects must produce non-zero vectors"
        assert float(np.sqrt((a.motion.astype(np.float64) ** 2).sum(axis=2)).max()) <= 96.0 + 1e-3   # max_vel clamp
    else:
        # first frame: prev_model = model and prev_viewproj = viewproj; prev_model * inverse(model) is the identity only up to rounding
        assert float(np.abs(a.motion).max()) < 1e-2


def _uniforms(sd, model, mat):
    u = capi.Uniforms()
    capi.set_f(u.model, model)
    capi.set_f(u.viewproj, sd.viewproj)
    capi.set_f(u.light_dir_ws, sd.sun_dir); capi.set_f(u.light_color, sd.sun_color); u.light_intensity = sd.sun_intensity
    capi.set_f(u.camera_pos, sd.cam_pos)
    capi.set_f(u.base_color, mat["base_color"]); u.metallic = mat["metallic"]; u.roughness = mat["roughness"]; u.ao = 1.0
    u.base_color_tex = mat.get("tex", 0)
    return u


@pytest.mark.parametrize("depth", [True, False])
@pytest.mark.parametrize("near_clip", [False, True])
def test_rasterize_mesh_aovs_through_reference_callbacks(port, reference, depth, near_clip):
    """Triangle ids / fragment counts: the reference side gets them through its own VS/FS callback API."""
    sd = scenes.scene_small(w=150, h=110, near_clip=near_clip, tex=True)
    A = HostAssets(sd.meshes, sd.textures)
    outs = []
    for o in (port, reference):
        hdr = np.zeros((sd.h, sd.w, 4), np.float32)
        dep = np.ones((sd.h, sd.w), np.float32) if depth else None
        tri = np.full((sd.h, sd.w), capi.TRI_ID_NONE, np.uint32)
        cov = np.zeros((sd.h, sd.w), np.uint32)
        tgt = o.make_target(sd.w, sd.h, hdr, dep, tri_id=tri, coverage=cov, zn=sd.zn, zf=sd.zf)
        key = 0
        stats = capi.Stats()
        for it in sd.items:
            model = o.model_from_transform(it["pos"], it.get("rot", (0, 0, 0)), it.get("scl", (1, 1, 1)))
            mat = it.get("material") or {"base_color": (0.8, 0.5, 0.2), "metallic": 0.1, "roughness": 0.5}
            st = o.rasterize_mesh(A, it["mesh"], capi.SHADER_PBR_MR, _uniforms(sd, model, mat), tgt, key_base=key)
            key += (len(sd.meshes[it["mesh"] - 1]["indices"]) // 3) * 8
            for k in ("tri_input", "tri_after_clip", "tri_raster"):
                setattr(stats, k, getattr(stats, k) + getattr(st, k))
        outs.append((hdr, dep, tri, cov, stats.as_dict()))
    (h0, d0, t0, c0, s0), (h1, d1, t1, c1, s1) = outs
    assert _same_bits(h0, h1)
    if depth:
        assert _same_bits(d0, d1)
    # the reference harness cannot see the fan index of a clipped polygon: compare (item, triangle) = key >> 3
    assert np.array_equal(t0 >> 3, t1 >> 3)
    if not depth:
        # painter mode: the reference shades every fragment that passes coverage + 1/w, so its FS-invocation count
        # equals the restatement's per-pixel coverage count
        assert np.array_equal(c0, c1)
    else:
        assert np.all(c1 <= c0) and np.array_equal(c1 > 0, c0 > 0)
    for k in ("tri_input", "tri_after_clip", "tri_raster"):
        assert s0[k] == s1[k]


def test_host_helpers_bit_exact(port, reference):
    from leisure_software_renderer_b200 import build, renderer
    build.build()
    rng = np.random.default_rng(7)
    for _ in range(50):
        pos, rot, scl = rng.uniform(-20, 20, 3), rng.uniform(-7, 7, 3), rng.uniform(0.05, 5, 3)
        a, b, c = port.model_from_transform(pos, rot, scl), reference.model_from_transform(pos, rot, scl), renderer.model_from_transform(pos, rot, scl)
        assert _same_bits(a, b) and _same_bits(a, c)
        eye, tgt = rng.uniform(-30, 30, 3), rng.uniform(-5, 5, 3)
        args = (eye, tgt, (0, 1, 0), float(rng.uniform(0.3, 2.0)), float(rng.uniform(0.5, 2.5)), 0.1, float(rng.uniform(50, 2000)))
        a, b, c = port.camera_viewproj(*args), reference.camera_viewproj(*args), renderer.camera_viewproj(*args)
        assert _same_bits(a, b) and _same_bits(a, c)


def test_light_packers_bit_exact(port, reference):
    rng = np.random.default_rng(11)
    for i in range(40):
        pos, color = rng.uniform(-10, 10, 3), rng.uniform(-0.2, 1.5, 3)
        rg, inten = float(rng.uniform(-1, 8)), float(rng.uniform(-1, 5))
        for jolt in (False, True):
            a = port.pack_point_light(pos, rg, color, inten, model=i % 3, power=1.25, bias=0.05, cutoff=0.0, jolt_bounds=jolt)
            b = reference.pack_point_light(pos, rg, color, inten, model=i % 3, power=1.25, bias=0.05, cutoff=0.0, jolt_bounds=jolt)
            assert np.array_equal(a, b), ("point", i, jolt)
        d = rng.uniform(-1, 1, 3)
        a = port.pack_spot_light(pos, rg, color, inten, d, float(rng.uniform(0, 1.2)), float(rng.uniform(0, 1.6)), model=i % 3, power=1.3)
        b = reference.pack_spot_light(pos, rg, color, inten, d, float(rng.uniform(0, 1.2)), float(rng.uniform(0, 1.6)), model=i % 3, power=1.3)
        # same rng draws for a and b are required: redo with fixed angles
        inner, outer = float(rng.uniform(0, 1.2)), float(rng.uniform(0, 1.6))
        a = port.pack_spot_light(pos, rg, color, inten, d, inner, outer, model=i % 3, power=1.3)
        b = reference.pack_spot_light(pos, rg, color, inten, d, inner, outer, model=i % 3, power=1.3)
        assert np.array_equal(a, b), ("spot", i)


def test_scene_light_records_match_reference_packer(reference):
    """scenes.pack_lights (numpy) lays records out exactly like make_point_culling_light (cos() of spots aside)."""
    lt = scenes.make_lights(16, 0, (-5, 0, -5), (5, 3, 5), seed=2)
    for r in lt:
        ref = reference.pack_point_light(r["position_range"][:3], float(r["position_range"][3]), r["color_intensity"][:3],
                                         float(r["color_intensity"][3]), model=1, power=1.25)
        assert np.array_equal(np.frombuffer(r.tobytes(), np.uint8), ref)


def test_tonemap_bit_exact(port, reference):
    rng = np.random.default_rng(3)
    hdr = rng.uniform(-0.5, 6.0, (37, 53, 4)).astype(np.float32)
    hdr[0, 0, :3] = (0.0, 1e-8, 1e6)
    for exposure, gamma in ((1.0, 2.2), (0.35, 1.0), (4.0, 2.4)):
        assert np.array_equal(port.pass_tonemap(hdr, exposure, gamma), reference.pass_tonemap(hdr, exposure, gamma))


def test_preserve_existing_depth(port, reference):
    """Quirk Q1: with preserve_existing_depth the lit pass tests LESS against the prepass depth (pass_pbr_forward.hpp:89-98)."""
    sd = scenes.scene_small(w=120, h=90)
    pre = harness.cpu_forward(port, sd, aov=False, tonemap=False)
    half = pre.depth.copy()
    half[:, : sd.w // 2] = 1.0  # left half: cleared depth -> surfaces pass; right half: equal depth -> everything rejected
    a = harness.cpu_forward(port, sd, aov=False, preserve_depth=True, init_depth=half)
    b = harness.cpu_forward(reference, sd, aov=False, preserve_depth=True, init_depth=half)
    assert _same_bits(a.hdr, b.hdr) and _same_bits(a.depth, b.depth)



def duplicate_id_pair():
    """Two frames in which two (and then three) items share one RenderItem::object_id: the reference keeps ONE previous model per key
    (Context::history.prev_model_by_object[key] = model, pass_pbr_forward.hpp:212: the last item with that key wins), so every item with
    the key gets that model as its prev_model in the next frame."""
    from leisure_software_renderer_b200 import scenes
    import dataclasses
    a = scenes.scene_small(w=176, h=110, motion=True, n_inst=4, seed=3)
    items = [dict(it) for it in a.items]
    for it, oid in zip(items[1:], (77, 77, 5, 77)):
        it["object_id"] = oid
    a = dataclasses.replace(a, items=items)
    b = a.moved(dpos=(0.3, 0.05, -0.2), drot=(0.0, 0.4, 0.1))
    return a, b


def last_duplicate_models(prev, cur, models):
    """prev_models per item of `cur` under the reference's map rule: the model of the LAST item of `prev` with the same key."""
    keys = [int(it.get("object_id", 0)) for it in prev.items]
    out = models.copy().reshape(len(keys), 16)
    last = {k: i for i, k in enumerate(keys) if k}
    for i, it in enumerate(cur.items):
        k = int(it.get("object_id", 0))
        if k in last:
            out[i] = models.reshape(len(keys), 16)[last[k]]
    return out


def test_duplicate_object_ids_share_the_last_previous_model(port, reference):
    prev, cur = duplicate_id_pair()
    pm = prev.models(port)
    ref = harness.cpu_forward(reference, cur, aov=False, motion=True, prev_models=pm)          # the reference's own map
    own = harness.cpu_forward(port, cur, aov=False, motion=True, prev_models=pm)               # every item its own previous model
    mapped = harness.cpu_forward(port, cur, aov=False, motion=True, prev_models=last_duplicate_models(prev, cur, pm))
    assert np.array_equal(mapped.motion.view(np.uint32), ref.motion.view(np.uint32))
    assert not np.array_equal(own.motion.view(np.uint32), ref.motion.view(np.uint32)), "the case must tell the two rules apart"
