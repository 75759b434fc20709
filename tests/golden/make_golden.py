"""Generates tests/golden/*.npz by running THE REFERENCE ITSELF (oracle/_ref/libshs_ref.so = the reference's own
headers compiled from /root/reference) on the deterministic scenes of tests/cases.py.  Run in the container that
has /root/reference; the committed fixtures let the GPU box (which has no reference tree) check both the oracle and
the CUDA path against reference outputs.  Light-list goldens come from the restatement (the reference's
light-culling headers need Jolt and cannot be compiled; 'parity unpinned' for that row)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))

import harness  # noqa: E402
import post_cases  # noqa: E402
from leisure_software_renderer_b200 import capi, scenes  # noqa: E402
from oracle.bindings import Oracle  # noqa: E402

GOLDEN = {
    "golden_pbr_96x72": (lambda: scenes.scene_small(w=96, h=72), {}),
    "golden_blinn_tex_clip_96x72": (lambda: scenes.scene_small(w=96, h=72, near_clip=True, tex=True, shading=capi.SHADING_BLINN), {}),
    "golden_painter_96x72": (lambda: scenes.scene_small(w=96, h=72, seed=4), {"depth": False}),
    "golden_shadow_pcf_96x72": (lambda: _shadow(), {"shadow": True}),
    "golden_c1_160x120": (lambda: scenes.scene_c1(160, 120), {}),
    "golden_sky_cubemap_96x72": (lambda: scenes.scene_small(w=96, h=72, sky="cubemap", tex=True), {}),
    "golden_sky_procedural_96x72": (lambda: scenes.scene_small(w=96, h=72, sky="procedural", seed=6), {}),
}

# motion-vector fixtures: (previous frame, current frame) -> reference output of the current frame with the previous one as history
GOLDEN_MOTION = {
    "golden_motion_96x72": lambda: _motion(),
}


def _motion():
    a = scenes.scene_small(w=96, h=72, motion=True, tex=True, n_inst=4, seed=2)
    return a, a.moved(cam_pos=(0.2, 4.05, -7.9))


def _shadow():
    sd = scenes.scene_small(w=96, h=72, tex=True)
    sd.fp.shadow_enable = 1
    sd.shadow_size = 128
    return sd


def main():
    ref = Oracle("reference")
    port = Oracle("port")
    for name, (make, kw) in GOLDEN.items():
        sd = make()
        f = harness.cpu_forward(ref, sd, aov=False, **kw)
        out = {"hdr": f.hdr, "ldr": f.ldr, "stats": np.array([f.stats[k] for k in ("tri_input", "tri_after_clip", "tri_raster")], np.uint64)}
        if f.depth is not None:
            out["depth"] = f.depth
        if f.shadow is not None:
            out["shadow"], out["lvp"] = f.shadow, f.lvp
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, {k: v.shape for k, v in out.items()})
    for name, make in GOLDEN_MOTION.items():
        prev, cur = make()
        f = harness.cpu_forward(ref, cur, aov=False, motion=True, prev_models=prev.models(ref))
        np.savez_compressed(os.path.join(HERE, name + ".npz"), hdr=f.hdr, ldr=f.ldr, depth=f.depth, motion=f.motion)
        print(name, f.motion.shape, float(np.abs(f.motion).max()))
    # depth-range / clustered bins: restatement only (the reference's headers need Jolt) -> "_port" fixture
    import test_light_bins_cpu as lb
    sd = lb.bins_scene()
    f = harness.cpu_forward(port, sd, aov=False)
    lo, hi = port.tile_depth_range(f.depth, 16, sd.zn, sd.zf)
    rng = np.random.default_rng(3)
    n = lo.size
    r01_min = (rng.random(n, dtype=np.float32) * np.float32(0.6)).astype(np.float32)
    r01_max = (r01_min + rng.random(n, dtype=np.float32) * np.float32(0.5)).astype(np.float32)
    r01_max[::7] = r01_min[::7]          # zero-thickness cells: NaN side planes keep every light touching the slab (classify_* never reject on NaN)
    bins = {"range_min": lo, "range_max": hi, "range01_min": r01_min, "range01_max": r01_max}
    for name, d in lb.descs(sd, mx=32).items():
        r = (lo, hi) if name == "view_depth" else ((r01_min, r01_max) if name == "depth01" else (None, None))
        c, i = port.light_cull_ex(sd.lights, d, *r)
        bins[name + "_counts"], bins[name + "_indices"] = c, i
    np.savez_compressed(os.path.join(HERE, "golden_light_bins_port.npz"), **bins)
    print("light bins", {k: int(v.sum()) for k, v in bins.items() if k.endswith("_counts")})
    # mixed light set (point x 3 attenuation models, spot, rect area, tube area, disabled, zero-intensity, zero-range):
    # the RECORDS come from the reference's own packers; how area lights shade is the oracle's (GLSL-derived) definition
    recs = []
    rng = np.random.default_rng(17)
    def P(lo, hi):
        return rng.uniform(lo, hi, 3).astype(np.float32)
    for k in range(60):
        pos = P((-5, 0.3, -5), (5, 2.5, 5)); col = rng.uniform(0.2, 1.0, 3).astype(np.float32)
        kind = k % 6
        model, power, bias, cutoff = k % 3, float(rng.uniform(0.5, 2.5)), float(rng.uniform(0.01, 0.3)), float(rng.uniform(0.0, 0.05) if k % 4 == 0 else 0.0)
        flags = 7 if k % 11 else 6          # every 11th light has LightFlagEnabled cleared
        inten = 0.0 if k == 7 else float(rng.uniform(1.5, 4.0))
        rad = 0.0 if k == 13 else float(rng.uniform(2.0, 5.0))
        if kind in (0, 1, 2):
            r = ref.pack_point_light(pos, rad, col, inten, model=model, power=power, bias=bias, cutoff=cutoff, jolt_bounds=(kind == 2))
            if flags != 7:
                r.view(np.uint32)[26] = flags   # type_shape_flags.z
        elif kind == 3:
            d = P((-1, -2, -1), (1, -0.5, 1))
            r = ref.pack_spot_light(pos, rad, col, inten, d, float(rng.uniform(0.1, 0.5)), float(rng.uniform(0.3, 0.9)), model=model, power=power, bias=bias, cutoff=cutoff)
        elif kind == 4:
            d = P((-0.4, -1, -0.4), (0.4, -0.6, 0.4)); right = P((0.6, -0.2, -0.4), (1.0, 0.2, 0.4))
            r = ref.pack_rect_light(pos, rad, col, inten, d, right, float(rng.uniform(0.3, 1.5)), float(rng.uniform(0.2, 1.0)), flags=flags, model=model, power=power, bias=bias, cutoff=cutoff)
        else:
            ax = P((-1, -0.3, -1), (1, 0.3, 1))
            r = ref.pack_tube_light(pos, rad, col, inten, ax, float(rng.uniform(0.3, 2.0)), float(rng.uniform(0.05, 0.4)), flags=flags, model=model, power=power, bias=bias, cutoff=cutoff)
        recs.append(r)
    np.savez_compressed(os.path.join(HERE, "golden_area_lights.npz"), records=np.stack(recs))
    print("mixed light records", len(recs))
    post = {}
    for name, (make, p) in post_cases.blur_cases().items():
        ldr, depth, motion = make()
        post["blur_" + name] = ref.pass_motion_blur(p, ldr, motion, depth)
    for name, (make, p, with_depth) in post_cases.shafts_cases().items():
        ldr, depth, _ = make()
        post["shafts_" + name] = ref.pass_light_shafts(p, ldr, depth if with_depth else None)
    np.savez_compressed(os.path.join(HERE, "golden_post_passes.npz"), **post)
    print("post passes", sorted(post))
    sd = scenes.scene_small(w=320, h=200, lights=64)
    counts, indices = port.light_cull(sd.lights, sd.viewproj, sd.w, sd.h, 16, 128)
    np.savez_compressed(os.path.join(HERE, "golden_light_lists_320x200_port.npz"), counts=counts, indices=indices)
    print("light lists", counts.shape, int(counts.max()))


if __name__ == "__main__":
    main()
