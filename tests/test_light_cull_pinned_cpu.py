"""Tile / depth-range / clustered light lists PINNED against the reference's own
builders: shs/lighting/jolt_light_culling.hpp:135-412 (with geometry/jolt_culling.hpp, frustum_culling.hpp, scene_shape.hpp,
jolt_adapter.hpp) compiled where it lies with SHS_HAS_JOLT=1 against a JoltPhysics DECLARATION shim (oracle/jolt_shim: the headers
use Jolt only to fetch a light's world bounds) by oracle/ref_lightcull_harness.cpp.  The restatement (oracle.cpp light_cull_bins)
must produce the same per-bin counts and the same index lists, bit for bit, for all four builders.  A light enters the reference as
its world AABB; the sphere / AABB the reference derives from it (SceneShape::bounding_sphere / world_aabb, after the LH <-> RH round
trip) are written into the CullingLightGPU records the restatement reads, so both sides classify the same numbers."""
import os

import numpy as np
import pytest

import fuzz_cases
from leisure_software_renderer_b200 import capi, scenes
from oracle.bindings import LightCullReference


@pytest.fixture(scope="module")
def ref():
    if not LightCullReference.available() and not os.path.isdir("/root/reference"):
        pytest.skip("oracle/_ref/libshs_lightcull_ref.so not built and /root/reference absent")
    return LightCullReference()


def with_reference_bounds(ref, lights):
    """The records with the bounds the reference derives from each light's AABB (returns records, aabbs)."""
    r = np.array(lights, copy=True)
    aabbs = np.concatenate([r["cull_aabb_min"][:, :3], r["cull_aabb_max"][:, :3]], axis=1).astype(np.float32)
    b = ref.bounds(aabbs)
    r["cull_sphere"][:, :] = b[:, 0:4]
    r["cull_aabb_min"][:, :3] = b[:, 4:7]
    r["cull_aabb_max"][:, :3] = b[:, 7:10]
    return r, aabbs


def test_reference_bounds_are_the_jolt_sphere_of_the_aabb(ref):
    """bounding sphere = centre of the AABB, radius = length of its half extent; the AABB survives the
    SHS -> Jolt -> SHS round trip unchanged."""
    lights = scenes.make_lights(40, 10, (-9, -2, -9), (9, 4, 9), seed=3, range_lo=0.2, range_hi=6.0)
    r, aabbs = with_reference_bounds(ref, lights)
    assert np.array_equal(r["cull_aabb_min"][:, :3], aabbs[:, :3]) and np.array_equal(r["cull_aabb_max"][:, :3], aabbs[:, 3:])
    half = np.float32(0.5) * (aabbs[:, 3:] - aabbs[:, :3])
    assert np.allclose(r["cull_sphere"][:, 3], np.sqrt((half.astype(np.float64) ** 2).sum(axis=1)), rtol=1e-6)
    assert np.array_equal(r["cull_sphere"][:, :3], np.float32(0.5) * (aabbs[:, :3] + aabbs[:, 3:]))


@pytest.mark.parametrize("seed", list(range(60)))
def test_fuzz_light_lists_equal_the_reference(port, ref, seed):
    lights, descs = fuzz_cases.light_bins(seed)
    r, aabbs = with_reference_bounds(ref, lights)
    for name, desc, lo, hi in descs:
        pc, pi = port.light_cull_ex(r, desc, lo, hi)
        rc, ri = ref.light_cull(aabbs, desc, lo, hi)
        assert np.array_equal(pc, rc), f"seed {seed} {name}: counts differ in {int(np.count_nonzero(pc != rc))} of {pc.size} bins"
        keep = np.arange(pi.shape[1])[None, :] < np.minimum(pc, pi.shape[1])[:, None]
        assert np.array_equal(pi[keep], ri[keep]), f"seed {seed} {name}: index lists differ"
    if seed == 0:
        pc, pi = port.light_cull(r, np.frombuffer(bytes(descs[0][1].view_proj), dtype=np.float32), descs[0][1].viewport_w, descs[0][1].viewport_h,
                                 descs[0][1].tile_size, descs[0][1].max_per_bin)
        rc, ri = ref.light_cull(aabbs, descs[0][1])
        assert np.array_equal(pc, rc)


def test_c2_shaped_light_lists_equal_the_reference(port, ref):
    """The bench workload's shape at a reduced size: 1024 point / spot lights, 16-px tiles, cap 128, 480x270."""
    sd = scenes.scene_small(w=480, h=270, lights=1024, seed=2)
    r, aabbs = with_reference_bounds(ref, sd.lights)
    desc = capi.LightCullDesc(sd.viewproj, sd.w, sd.h, capi.LIGHT_CULL_TILED, 16, 128, z_near=sd.zn, z_far=sd.zf)
    pc, pi = port.light_cull_ex(r, desc)
    rc, ri = ref.light_cull(aabbs, desc)
    assert np.array_equal(pc, rc) and int(pc.sum()) > 1000
    keep = np.arange(128)[None, :] < np.minimum(pc, 128)[:, None]
    assert np.array_equal(pi[keep], ri[keep])


def test_c2_full_size_light_lists_equal_the_reference(port, ref):
    """BASELINE configs[1] at its full size: 1920x1080, 16-px tiles (8160 tiles), 1024 point / spot lights, cap 128 -- about 548 k list
    entries, up to ~300 lights in a tile (so the cap cuts lists: counts stay uncapped on both sides, the first 128 entries agree)."""
    sd = scenes.scene_c2()
    r, aabbs = with_reference_bounds(ref, sd.lights)
    desc = capi.LightCullDesc(sd.viewproj, sd.w, sd.h, capi.LIGHT_CULL_TILED, 16, 128, z_near=sd.zn, z_far=sd.zf)
    pc, pi = port.light_cull_ex(r, desc)
    rc, ri = ref.light_cull(aabbs, desc)
    assert np.array_equal(pc, rc) and int(pc.sum()) > 400000 and int(pc.max()) > 128
    keep = np.arange(128)[None, :] < np.minimum(pc, 128)[:, None]
    assert np.array_equal(pi[keep], ri[keep])
