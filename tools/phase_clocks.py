"""Per-phase cycles of tile_kernel CTAs (debug build with -DSHSB_PHASE_CLOCKS, SHSB_LIB=...libshsb_clk.so):
python tools/phase_clocks.py [c2|c4]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from leisure_software_renderer_b200 import capi, scenes
from leisure_software_renderer_b200.renderer import Context

which = sys.argv[1] if len(sys.argv) > 1 else "c2"
sd = scenes.scene_c2() if which == "c2" else scenes.scene_c4()
ctx = Context(0)
for m in sd.meshes:
    ctx.mesh_upload(m["positions"], m["normals"], m["uvs"], m["indices"])
for t in sd.textures:
    ctx.texture_upload(t)
if sd.lights is not None:
    ctx.lights_upload(sd.lights.view(np.uint8))
hdr = ctx.rt_create(capi.RT_COLOR_HDR, sd.w, sd.h); dm = ctx.rt_create(capi.RT_DEPTH_MOTION, sd.w, sd.h, sd.zn, sd.zf); ldr = ctx.rt_create(capi.RT_COLOR_LDR, sd.w, sd.h)
for _ in range(3):
    ctx.frame_forward_plus(sd.scene, sd.fp, hdr, dm, ldr)
out = (C.c_ulonglong * 32)()
ctx.lib.shsb_debug_phase_clocks(out, 1)
N = 10
if len(sys.argv) > 2 and sys.argv[2] == "d2h":
    # a back-to-back 8.3 MB device -> pinned-host copy loop on another stream while the frames run
    import torch
    copy = torch.cuda.Stream()
    d_small = torch.empty(1920 * 1080 * 4, dtype=torch.uint8, device="cuda")
    h_small = [torch.empty(1920 * 1080 * 4, dtype=torch.uint8).pin_memory() for _ in range(2)]
    with torch.cuda.stream(copy):
        for i in range(200):
            h_small[i % 2].copy_(d_small, non_blocking=True)
    print("with a concurrent D2H copy loop")
for _ in range(N):
    ctx.frame_forward_plus(sd.scene, sd.fp, hdr, dm, ldr)
ctx.lib.shsb_debug_phase_clocks(out, 1)
a = np.array(list(out), dtype=np.float64).reshape(4, 8)
names = ["raster", "resolve+phaseA", "barrier+stats", "light staging", "light loop", "phase C", "CTAs", "prologue"]
for cls in range(3):
    n = a[cls, 6]
    if n == 0:
        continue
    tot = a[cls, [7, 0, 1, 2, 3, 4, 5]].sum()
    print(f"class {cls}: {int(n / N)} tiles/frame, mean {tot / n / 1965:.2f} us per CTA (sum over class = {tot / N / 1965 / 592:.1f} us of a 592-slot machine)")
    for i in [7, 0, 1, 2, 3, 4, 5]:
        print(f"    {names[i]:16s} {a[cls, i] / n:9.0f} cycles  {a[cls, i] / tot * 100:5.1f}%")
