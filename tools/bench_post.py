"""Device time and achieved bandwidth of the post-pass kernels at a BASELINE frame size (default 1080p).
Usage (GPU box): python tools/bench_post.py [w h] ; prints one JSON line per pass.  Algorithmic bytes per pixel
(DESIGN.md section 4.3): motion blur 4 (LDR in) + 8 (motion) + 4 (depth) + 4 (LDR out) = 20; light shafts 4 + 4 (depth) + 4 (out)
= 12 (+ 8 for the luma plane written and read once); TAA 4 + 4 in, 4 + 4 out = 16."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import post_cases  # noqa: E402
from leisure_software_renderer_b200 import capi  # noqa: E402
from leisure_software_renderer_b200.renderer import Context  # noqa: E402


def main():
    w, h = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) >= 3 else (1920, 1080)
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    gpu = Context(0)
    stream = torch.cuda.ExternalStream(gpu.stream())
    # 6 rotating target sets so that the working set (6 x 20 B/px at 1080p = 250 MB) exceeds the 126 MB L2
    sets = []
    for i in range(6):
        ldr, depth, motion = post_cases.planes(w, h, 30 + i, max_motion=40.0)
        src = gpu.rt_create(capi.RT_COLOR_LDR, w, h); dst = gpu.rt_create(capi.RT_COLOR_LDR, w, h)
        dm = gpu.rt_create(capi.RT_DEPTH_MOTION, w, h, 0.1, 1000.0)
        gpu.rt_upload(src, capi.PLANE_COLOR, ldr); gpu.rt_upload(dm, capi.PLANE_DEPTH, depth); gpu.rt_upload(dm, capi.PLANE_MOTION, motion)
        sets.append((src, dst, dm))
    mb = capi.MotionBlurParams()
    _, sp, _ = post_cases.shafts_cases()["sun_in_view"]
    passes = {
        "motion_blur": (lambda s: gpu.pass_motion_blur(mb, s[0], s[1], s[2]), 20),
        "light_shafts": (lambda s: gpu.pass_light_shafts(sp, s[0], s[1], s[2]), 12),
        "taa": (lambda s: gpu.pass_taa(s[1]), 16),
    }
    for name, (fn, bpp) in passes.items():
        for s in sets:
            fn(s)
        gpu.sync()
        n = 60
        with torch.cuda.stream(stream):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for i in range(n):
                fn(sets[i % len(sets)])
            e1.record(stream)
        gpu.sync()
        ms = e0.elapsed_time(e1) / n
        gbs = w * h * bpp / (ms * 1e-3) / 1e9
        print(json.dumps({"pass": name, "w": w, "h": h, "ms": ms, "algorithmic_bytes": w * h * bpp, "achieved_gbs": gbs, "frac_of_measured_hbm_peak": gbs / peak}))


def bins():
    """Depth reduce + the depth-range / clustered bin builders on the C2 scene (1080p, 1024 lights)."""
    from leisure_software_renderer_b200 import scenes
    sd = scenes.scene_c2()
    gpu = Context(0)
    for m in sd.meshes:
        gpu.mesh_upload(m["positions"], m["normals"], m["uvs"], m["indices"])
    gpu.lights_upload(sd.lights.view(np.uint8))
    hdr = gpu.rt_create(capi.RT_COLOR_HDR, sd.w, sd.h); dm = gpu.rt_create(capi.RT_DEPTH_MOTION, sd.w, sd.h, sd.zn, sd.zf)
    fp = capi.FrameParams.from_buffer_copy(sd.fp)
    fp.light_culling = 0
    gpu.pass_pbr_forward(sd.scene, fp, hdr, dm)
    stream = torch.cuda.ExternalStream(gpu.stream())
    import ctypes as C
    mk = lambda mode, **kw: capi.LightCullDesc(sd.viewproj, sd.w, sd.h, mode, 16, 128, z_near=sd.zn, z_far=sd.zf, **kw)
    descs = {"tiled (two-level, shsb_light_cull)": mk(capi.LIGHT_CULL_TILED), "tiled view-depth range": mk(capi.LIGHT_CULL_TILED_VIEW_DEPTH),
             "clustered x16": mk(capi.LIGHT_CULL_CLUSTERED, depth_slices=16)}
    calls = {"tile_depth_range": lambda: gpu.lib.shsb_tile_depth_range(gpu.h, dm, 16)}
    gpu.lib.shsb_tile_depth_range(gpu.h, dm, 16)
    for name, d in descs.items():
        calls["light_cull_ex " + name] = (lambda d=d: gpu.lib.shsb_light_cull_ex(gpu.h, C.byref(d), None, None))
    for name, fn in calls.items():
        for _ in range(3):
            assert fn() == 0
        gpu.sync()
        n = 30
        with torch.cuda.stream(stream):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(n):
                fn()
            e1.record(stream)
        gpu.sync()
        print(json.dumps({"call": name, "w": sd.w, "h": sd.h, "lights": len(sd.lights), "ms": e0.elapsed_time(e1) / n}))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "bins":
        bins()
    else:
        main()
