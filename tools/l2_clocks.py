"""Per-CTA clocks of legacy2_raster_kernel (debug build libshsb_clk.so: python -m leisure_software_renderer_b200.build --phase-clocks;
run with SHSB_LIB=.../libshsb_clk.so): renders the shipped soft-shadow frame of tools/bench_legacy2.py once and prints, for the LAST lit
draw, the distribution over CTAs of the cycles spent in the staging / visit loop and in the final shading."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from bench_legacy2 import shipped_scene, FLT_MAX  # noqa: E402
from test_zz_gpu_legacy2 import fill_uniforms  # noqa: E402
from leisure_software_renderer_b200 import capi  # noqa: E402
from leisure_software_renderer_b200.renderer import Context  # noqa: E402


def main():
    gpu = Context(0)
    sc = shipped_scene()
    W, H, sm = sc["W"], sc["H"], sc["sm"]
    meshes = [gpu.mesh_upload(pos, nrm, uv, None) for pos, nrm, uv, *_ in sc["objs"]]
    tex = gpu.texture_upload(sc["texture"])
    c, z, s = gpu.rt_create(capi.RT_COLOR_LDR, W, H), gpu.rt_create(capi.RT_DEPTH_MOTION, W, H), gpu.rt_create(capi.RT_SHADOW, sm, sm)
    gpu.rt_clear(s, capi.PLANE_DEPTH, FLT_MAX)
    for m, o in zip(meshes, sc["objs"]):
        gpu.legacy2_shadow_draw(m, sc["f32"](o[3]), sc["light_vp"], s, 160, 160)
    gpu.rt_clear(z, capi.PLANE_DEPTH, FLT_MAX)
    tx, ty = (W + 15) // 16, (H + 15) // 16
    for k, (m, o) in enumerate(zip(meshes, sc["objs"])):
        u = fill_uniforms(sc, o[3], o[4], o[5], tex, prev=False)
        gpu.legacy2_draw_softshadow(m, u, s, c, z)
        gpu.sync()
        out = np.zeros((2, 8192), np.uint64)
        assert gpu.lib.shsb_debug_l2_clocks(out.ctypes.data_as(C.c_void_p)) == 0
        for name, a in (("staging+visit", out[0][:tx * ty]), ("shading", out[1][:tx * ty])):
            a = a.astype(np.float64) / 1965.0  # us at 1965 MHz
            q = np.percentile(a, [50, 90, 99, 100])
            print(f"draw {k} ({len(o[0]) // 3} tris) {name:14s}: mean {a.mean():8.2f} us  p50 {q[0]:8.2f}  p90 {q[1]:8.2f}  p99 {q[2]:8.2f}  max {q[3]:8.2f}  sum {a.sum() / 1e3:8.2f} ms")
        heavy = np.argsort(out[0][:tx * ty] + out[1][:tx * ty])[-5:]
        print("   heaviest tiles (x, y):", [(int(i % tx), int(i // tx)) for i in heavy])


if __name__ == "__main__":
    main()
