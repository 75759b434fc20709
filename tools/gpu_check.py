"""Diagnostic (not a test): runs several scenes on the GPU and on the CPU oracle and prints mismatch statistics."""
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np

import harness
from leisure_software_renderer_b200 import capi, scenes
from leisure_software_renderer_b200.renderer import Context
from oracle.bindings import Oracle


def report(name, g, c, depth=True):
    out = [name]
    for k in ("tri_input", "tri_after_clip", "tri_raster", "frag_covered"):
        out.append(f"{k}={g.stats[k]}/{c.stats[k]}")
    if g.tri_id is not None and c.tri_id is not None:
        out.append(f"mask_diff={int(np.count_nonzero((g.tri_id != capi.TRI_ID_NONE) != (c.tri_id != capi.TRI_ID_NONE)))}")
        out.append(f"id_diff={int(np.count_nonzero(g.tri_id != c.tri_id))}")
        out.append(f"cov_diff={int(np.count_nonzero(g.coverage != c.coverage))}")
    if depth and g.depth is not None:
        u = harness.ulp_diff(g.depth, c.depth)
        out.append(f"depth_ulp_max={int(u.max())} n>0={int(np.count_nonzero(u))}")
    if c.counts is not None and g.counts is not None:
        out.append(f"count_diff={int(np.count_nonzero(g.counts != c.counts))} max_count={int(c.counts.max())}")
        m = g.indices.shape[1]
        valid = np.arange(m)[None, :] < np.minimum(c.counts, m)[:, None]
        out.append(f"list_diff={int(np.count_nonzero(g.indices[valid] != c.indices[valid]))}")
    if c.shadow is not None and g.shadow is not None:
        u = harness.ulp_diff(g.shadow, c.shadow)
        out.append(f"shadow_ulp_max={int(u.max())} n>0={int(np.count_nonzero(u))} lvp_equal={np.array_equal(g.lvp, c.lvp)}")
    out.append(f"psnr={harness.psnr(g.hdr[..., :3], c.hdr[..., :3]):.1f}")
    d = np.abs(g.ldr.astype(np.int32) - c.ldr.astype(np.int32))
    out.append(f"ldr_max={int(d.max())} n_off={int(np.count_nonzero(d.max(axis=2)))}")
    print("  ".join(out), flush=True)


def main():
    ctx = Context(0)
    port = Oracle("port")
    cases = [
        ("c1", scenes.scene_c1(), {}),
        ("small_pbr", scenes.scene_small(), {}),
        ("small_blinn_tex", scenes.scene_small(tex=True, shading=capi.SHADING_BLINN), {}),
        ("nearclip", scenes.scene_small(near_clip=True, tex=True), {}),
        ("painter", scenes.scene_small(), {"depth": False}),
        ("shadow", scenes.scene_small(w=320, h=200), {"shadow": True}),
        ("fplus", scenes.scene_small(w=320, h=200, lights=64), {"forward_plus": True}),
        ("fplus_fused", scenes.scene_small(w=333, h=207, lights=64, tex=True), {"forward_plus": True, "fused": True}),
        ("c2_small", scenes.scene_c2(w=640, h=360, grid=4, n_point=96, n_spot=32), {"forward_plus": True, "fused": True}),
    ]
    for name, sd, kw in cases:
        try:
            if name == "shadow":
                sd.fp.shadow_enable = 1
            t0 = time.time()
            g = harness.gpu_forward(ctx, sd, **kw)
            t1 = time.time()
            ckw = {k: v for k, v in kw.items() if k != "fused"}
            c = harness.cpu_forward(port, sd, **ckw)
            t2 = time.time()
            report(f"{name} [{sd.w}x{sd.h} gpu {t1 - t0:.3f}s cpu {t2 - t1:.3f}s]", g, c, depth=kw.get("depth", True))
        except Exception:
            print(name, "FAILED")
            traceback.print_exc()
    print("launches", ctx.launch_count(), "stage ms", ctx.last_stage_ms())


if __name__ == "__main__":
    main()
