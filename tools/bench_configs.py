"""Times every BASELINE.json config at full size on one GPU (secondary to bench.py, which is the C2 headline):
python tools/bench_configs.py [c1 c2 c3 c4 c5] [--iters N] -> one JSON line per config.

Per config: the passes the reference would run for it (PassShadowMap -> PassPBRForward (+Forward+ cull) -> PassTonemap),
wall-clock per frame with a device synchronise on both sides (frames here are 0.1-30 ms), the library's per-stage CUDA-event
times, rasterizer statistics, algorithmic bytes per SURVEY.md 8(d) and the fraction of the measured HBM peak they imply.
"""
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np

from leisure_software_renderer_b200 import capi, scenes
from leisure_software_renderer_b200.renderer import Context


def hbm_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def run(name, sd, iters, shadow=False, forward_plus=False):
    ctx = Context(0)
    for m in sd.meshes:
        ctx.mesh_upload(m["positions"], m["normals"], m["uvs"], m["indices"])
    for t in sd.textures:
        ctx.texture_upload(t)
    if sd.lights is not None:
        ctx.lights_upload(sd.lights.view(np.uint8))
    hdr = ctx.rt_create(capi.RT_COLOR_HDR, sd.w, sd.h)
    dm = ctx.rt_create(capi.RT_DEPTH_MOTION, sd.w, sd.h, sd.zn, sd.zf)
    ldr = ctx.rt_create(capi.RT_COLOR_LDR, sd.w, sd.h)
    sh = ctx.rt_create(capi.RT_SHADOW, sd.shadow_size, sd.shadow_size) if shadow else 0
    fp = capi.FrameParams.from_buffer_copy(sd.fp)
    fp.light_culling = 1 if forward_plus else 0
    fp.shadow_enable = 1 if shadow else 0

    def frame(stats=False):
        out = {}
        if shadow:
            lvp = ctx.pass_shadow_map(sd.scene, fp, sh)
            out["shadow_ms"] = float(ctx.last_stage_ms()[5]) if stats else 0.0
            st = ctx.pass_pbr_forward(sd.scene, fp, hdr, dm, sh, lvp, want_stats=stats)
            ctx.pass_tonemap(hdr, ldr, fp.exposure, fp.gamma)
        else:
            st = ctx.frame_forward_plus(sd.scene, fp, hdr, dm, ldr, want_stats=stats)
        if stats:
            ms = ctx.last_stage_ms()
            out.update({"vertex_clip_setup_ms": float(ms[0]), "binning_ms": float(ms[1]), "tile_ms": float(ms[2]), "light_cull_ms": float(ms[3]), "front_to_tile_end_ms": float(ms[5]), "stats": st.as_dict()})
        return out

    info = frame(stats=True)   # sizes the arenas
    frame(); frame()
    ctx.sync()
    times = []
    for _ in range(iters):
        ctx.sync(); t0 = time.perf_counter()
        frame()
        ctx.sync(); times.append((time.perf_counter() - t0) * 1e3)
    # determinism + sanity properties at full size
    a = ctx.rt_download(ldr)
    frame(); ctx.sync()
    b = ctx.rt_download(ldr)
    dep = ctx.rt_download(dm, capi.PLANE_DEPTH)
    st = info["stats"]
    props = {"deterministic": bool(np.array_equal(a, b)), "depth_in_0_1": bool(dep.min() >= 0.0 and dep.max() <= 1.0),
             "covered_pixels": int(np.count_nonzero(dep < 1.0)), "frag_shaded": int(st["frag_shaded"])}
    ms = statistics.median(times)
    px = sd.w * sd.h
    b_alg = px * 24 + (2 * 4 * sd.shadow_size ** 2 if shadow else 0) + (0 if sd.lights is None else 160 * len(sd.lights))
    b_alg += sum(32 * len(m["positions"]) + 4 * len(m["indices"]) for m in sd.meshes) + 96 * len(sd.items) + sum(t.size for t in sd.textures)
    line = {"config": name, "scene": sd.name, "resolution": [sd.w, sd.h], "frame_ms": ms, "frame_ms_min": min(times), "frames_per_s": 1e3 / ms,
            "mtri_per_s": st["tri_input"] / ms / 1e3, "mfrag_per_s": st["frag_covered"] / ms / 1e3,
            "algorithmic_bytes": b_alg, "hbm_frac": b_alg / 1e9 / (ms / 1e3) / hbm_peak(), "iters": iters, **info, "properties": props,
            "timing": "wall clock with device synchronise on both sides of each frame (includes host submission)"}
    print(json.dumps(line), flush=True)
    ctx.close()


def main():
    args = [a for a in sys.argv[1:] if a in ("c1", "c2", "c3", "c4", "c5")]
    iters = int(sys.argv[sys.argv.index("--iters") + 1]) if "--iters" in sys.argv else 20
    which = args or ["c1", "c2", "c3", "c4", "c5"]
    if "c1" in which:
        run("C1", scenes.scene_c1(), iters)
    if "c2" in which:
        run("C2", scenes.scene_c2(), iters, forward_plus=True)
    if "c3" in which:
        run("C3", scenes.scene_c3(), iters, shadow=True)
    if "c4" in which:
        run("C4", scenes.scene_c4(), iters)
    if "c5" in which:
        run("C5", scenes.scene_c5(), max(3, iters // 4), forward_plus=True)


if __name__ == "__main__":
    main()
