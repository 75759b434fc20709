"""Which part of the e2e step limits the pipeline: frames only / + lights upload / + LDR read-back / both."""
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
N = 300
for up, down in ((0, 0), (1, 0), (0, 1), (1, 1)):
    for rep in range(2):
        ctx.timing_enable(True)
        ctx.sync(); t0 = time.perf_counter()
        for i in range(N):
            if up:
                ctx.lights_upload(lights.numpy())
            ctx.frame_forward_plus(sd.scene, sd.fp, *sets[i % 4], want_stats=False)
            if down:
                ctx.rt_download_async(sets[i % 4][2], capi.PLANE_COLOR, host[i % 2].data_ptr(), W * H * 4)
        ctx.sync(); dt = (time.perf_counter() - t0) / N * 1e6
        st = ctx.timing_collect().mean(axis=0) * 1e3
        ctx.timing_enable(False)
    print(f"lights upload={up} read-back={down}: {dt:7.1f} us/frame   stages us: geometry {st[0]:.1f} binning {st[1]:.1f} tile {st[2]:.1f} front-begin..tile-end {st[3]:.1f}")

# ---- the same read-back done with torch streams so that each D2H copy can be timed
main = torch.cuda.ExternalStream(ctx.stream(), device=0)
copy = torch.cuda.Stream(device=0)


def view(rt):
    ptr, nbytes = ctx.rt_device_ptr(rt, capi.PLANE_COLOR)

    class _Cai:
        __cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}
    return torch.as_tensor(_Cai(), device="cuda:0")


views = [view(s[2]) for s in sets]
evs = []
ctx.sync(); torch.cuda.synchronize(); t0 = time.perf_counter()
for i in range(N):
    ctx.frame_forward_plus(sd.scene, sd.fp, *sets[i % 4], want_stats=False)
    ev = torch.cuda.Event(); ev.record(main)
    copy.wait_event(ev)
    with torch.cuda.stream(copy):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(copy)
        host[i % 2].copy_(views[i % 4], non_blocking=True)
        e1.record(copy)
    evs.append((e0, e1))
torch.cuda.synchronize(); ctx.sync(); dt = (time.perf_counter() - t0) / N * 1e6
d = np.array([a.elapsed_time(b) for a, b in evs]) * 1e3
gaps = np.array([evs[k][1].elapsed_time(evs[k + 1][0]) for k in range(N - 1)]) * 1e3
print(f"torch-stream read-back (no hazard tracking): {dt:.1f} us/frame; D2H copy mean {d.mean():.1f} us (min {d.min():.1f}, max {d.max():.1f}); gap between copies mean {gaps.mean():.1f} us")
