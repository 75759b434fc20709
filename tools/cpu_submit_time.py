"""How long does the host need to SUBMIT one C2 frame (no device sync inside the loop)?"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from leisure_software_renderer_b200 import capi, scenes
from leisure_software_renderer_b200.renderer import Context

ctx = Context(0)
sd = scenes.scene_c2()
for m in sd.meshes:
    ctx.mesh_upload(m["positions"], m["normals"], m["uvs"], m["indices"])
ctx.lights_upload(sd.lights)
sets = [(ctx.rt_create(capi.RT_COLOR_HDR, sd.w, sd.h), ctx.rt_create(capi.RT_DEPTH_MOTION, sd.w, sd.h, sd.zn, sd.zf), ctx.rt_create(capi.RT_COLOR_LDR, sd.w, sd.h)) for _ in range(4)]
for i in range(10):
    ctx.frame_forward_plus(sd.scene, sd.fp, *sets[i % 4], want_stats=False)
ctx.sync()
for n in (20, 200, 1000):
    t0 = time.perf_counter()
    for i in range(n):
        ctx.frame_forward_plus(sd.scene, sd.fp, *sets[i % 4], want_stats=False)
    t1 = time.perf_counter()
    ctx.sync()
    t2 = time.perf_counter()
    print(f"n={n}: submit {1e6 * (t1 - t0) / n:.1f} us/frame, total {1e6 * (t2 - t0) / n:.1f} us/frame")
# pure host cost: 2 frames after a sync never wait on the 3-slot staging ring
ts = []
for rep in range(50):
    ctx.sync()
    t0 = time.perf_counter()
    ctx.frame_forward_plus(sd.scene, sd.fp, *sets[0], want_stats=False)
    ctx.frame_forward_plus(sd.scene, sd.fp, *sets[1], want_stats=False)
    t1 = time.perf_counter()
    ts.append((t1 - t0) / 2)
print(f"host-only submit cost: median {1e6 * sorted(ts)[len(ts)//2]:.1f} us/frame, min {1e6 * min(ts):.1f}")

ctx.host_submit_us()
for i in range(200):
    ctx.frame_forward_plus(sd.scene, sd.fp, *sets[i % 4], want_stats=False)
ctx.sync()
h = ctx.host_submit_us()
print("per-frame host us: draw-list %.1f staging %.1f arena %.1f capture/enqueue %.1f update+launch %.1f" % tuple(h[:5] / h[5]))
