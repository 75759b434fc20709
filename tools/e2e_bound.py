"""What bounds the e2e frame period once stage events are off: the render chain or the read-back chain?
Frames with read-back at three render costs: full Forward+ (tile kernel ~140 us), sun only (no local lights: a much shorter
tile kernel), and no rendering at all (read-back only)."""
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
host = [torch.empty(W * H * 4, dtype=torch.uint8).pin_memory() for _ in range(2)]
ctx.frame_forward_plus(sd.scene, sd.fp, *sets[0])
fp_sun = capi.FrameParams.from_buffer_copy(sd.fp)
fp_sun.light_culling = 0
N = 400
for name, fp, render, down in (("Forward+ frames, no read-back", sd.fp, True, False), ("Forward+ frames + read-back", sd.fp, True, True),
                               ("sun-only frames, no read-back", fp_sun, True, False), ("sun-only frames + read-back", fp_sun, True, True),
                               ("read-back only", sd.fp, False, True)):
    for rep in range(2):
        ctx.sync(); t0 = time.perf_counter()
        for i in range(N):
            if render:
                ctx.lights_upload(lights.numpy())
                ctx.frame_forward_plus(sd.scene, fp, *sets[i % 4], want_stats=False)
            if down:
                ctx.rt_download_async(sets[i % 4][2], capi.PLANE_COLOR, host[i % 2].data_ptr(), W * H * 4)
        ctx.sync(); dt = (time.perf_counter() - t0) / N * 1e6
    print(f"{name:32s}: {dt:7.1f} us/frame")
