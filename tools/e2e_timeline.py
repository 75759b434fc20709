"""Timeline of the pipelined e2e loop (frames + LDR read-back): per frame, when the host submitted it, when its front
end ran, when its tile kernel ran and when its D2H copy ran -- to find what bounds the frame period."""
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
main = torch.cuda.ExternalStream(ctx.stream(), device=0)
copy = torch.cuda.Stream(device=0)


def view(rt):
    ptr, nbytes = ctx.rt_device_ptr(rt, capi.PLANE_COLOR)

    class _Cai:
        __cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}
    return torch.as_tensor(_Cai(), device="cuda:0")


views = [view(s[2]) for s in sets]
N = 60
for i in range(20):
    ctx.frame_forward_plus(sd.scene, sd.fp, *sets[i % 4], want_stats=False)
ctx.sync(); torch.cuda.synchronize()
ctx.timing_enable(True)
base = torch.cuda.Event(enable_timing=True); base.record(main)
t0 = time.perf_counter()
hs, evs = [], []
for i in range(N):
    hs.append((time.perf_counter() - t0) * 1e6)
    ctx.frame_forward_plus(sd.scene, sd.fp, *sets[i % 4], want_stats=False)
    ev = torch.cuda.Event(); ev.record(main)
    copy.wait_event(ev)
    with torch.cuda.stream(copy):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(copy)
        host[i % 2].copy_(views[i % 4], non_blocking=True)
        e1.record(copy)
    evs.append((e0, e1))
torch.cuda.synchronize(); ctx.sync()
tl = ctx.timing_collect_abs() * 1e3
print("frame | host submit | front begin  geom end  bin end | tile begin  tile end | copy begin  copy end   (us)")
for i in range(30, 46):
    c0, c1 = base.elapsed_time(evs[i][0]) * 1e3, base.elapsed_time(evs[i][1]) * 1e3
    f = tl[i]
    print(f"{i:5d} | {hs[i]:9.0f}   | {f[0]:9.0f} {f[1]:9.0f} {f[2]:9.0f} | {f[3]:9.0f} {f[4]:9.0f} | {c0:9.0f} {c1:9.0f}")

# ---- the library's own read-back path (hazard tracking, alternating copy streams): frame period only
for i in range(20):
    ctx.frame_forward_plus(sd.scene, sd.fp, *sets[i % 4], want_stats=False)
ctx.sync(); t0 = time.perf_counter()
for i in range(300):
    ctx.frame_forward_plus(sd.scene, sd.fp, *sets[i % 4], want_stats=False)
    ctx.rt_download_async(sets[i % 4][2], capi.PLANE_COLOR, host[i % 2].data_ptr(), W * H * 4)
ctx.sync()
print(f"library read-back path: {(time.perf_counter() - t0) / 300 * 1e6:.1f} us/frame")
