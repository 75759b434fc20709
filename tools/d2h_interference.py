"""Does a concurrent D2H copy slow the frame's kernels?  Device-resident frames (no read-back of the frames themselves)
with and without an UNRELATED 8.3 MB device -> pinned-host copy loop running on another stream.
Run twice: plain and with SHSB_NO_PIPELINE=1 (front end and tile kernel serialised on one stream)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from leisure_software_renderer_b200 import capi, scenes
from leisure_software_renderer_b200.renderer import Context

W, H = 1920, 1080
ctx = Context(0)
sd = scenes.scene_c2(W, H)
for m in sd.meshes:
    ctx.mesh_upload(m["positions"], m["normals"], m["uvs"], m["indices"])
lights = torch.from_numpy(np.ascontiguousarray(sd.lights).view(np.uint8).copy()).pin_memory()
ctx.lights_upload(lights.numpy())
sets = [(ctx.rt_create(capi.RT_COLOR_HDR, W, H), ctx.rt_create(capi.RT_DEPTH_MOTION, W, H, sd.zn, sd.zf), ctx.rt_create(capi.RT_COLOR_LDR, W, H)) for _ in range(4)]
ctx.frame_forward_plus(sd.scene, sd.fp, *sets[0])
copy = torch.cuda.Stream(device=0)
src = torch.empty(W * H * 4, dtype=torch.uint8, device="cuda")
dst = [torch.empty(W * H * 4, dtype=torch.uint8).pin_memory() for _ in range(2)]
big = torch.empty(W * H * 4, dtype=torch.uint8).pin_memory()
N = 300
print("pipeline" if not os.environ.get("SHSB_NO_PIPELINE") else "SHSB_NO_PIPELINE=1")
for mode in ("none", "d2h", "h2d"):
    for rep in range(2):
        ctx.timing_enable(True)
        ctx.sync(); torch.cuda.synchronize(); t0 = time.perf_counter()
        for i in range(N):
            ctx.frame_forward_plus(sd.scene, sd.fp, *sets[i % 4], want_stats=False)
            if mode != "none":
                with torch.cuda.stream(copy):
                    if mode == "d2h":
                        dst[i % 2].copy_(src, non_blocking=True)
                    else:
                        src.copy_(big, non_blocking=True)
        ctx.sync(); torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / N * 1e6
        st = ctx.timing_collect().mean(axis=0) * 1e3
        ctx.timing_enable(False)
    print(f"concurrent copy={mode:5s}: {dt:7.1f} us/frame   stages us: front-begin..geometry {st[0]:.1f} binning {st[1]:.1f} tile {st[2]:.1f}")
