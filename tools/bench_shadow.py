"""The C3 shadow pass alone (BASELINE configs[2]: 4096^2 shadow map, 1.03 M triangles): python tools/bench_shadow.py [iters]
-> one JSON line: per-stage CUDA-event times (vertex / set-up, binning, tile raster), wall clock per pass, algorithmic bytes
(4 B per texel written once + 32 B per unique vertex + 4 B per index + 96 B per instance) against the measured HBM peak."""
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np

from leisure_software_renderer_b200 import capi, scenes
from leisure_software_renderer_b200.renderer import Context


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    size = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    sd = scenes.scene_c3(shadow_size=size)
    ctx = Context(0)
    for m in sd.meshes:
        ctx.mesh_upload(m["positions"], m["normals"], m["uvs"], m["indices"])
    sh = ctx.rt_create(capi.RT_SHADOW, size, size)
    fp = capi.FrameParams.from_buffer_copy(sd.fp)
    fp.shadow_enable = 1
    for _ in range(5):
        ctx.pass_shadow_map(sd.scene, fp, sh)
    stage, wall = [], []
    for _ in range(iters):
        ctx.sync()
        t0 = time.perf_counter()
        ctx.pass_shadow_map(sd.scene, fp, sh)
        ctx.sync()
        wall.append((time.perf_counter() - t0) * 1e3)
        stage.append(ctx.last_stage_ms().copy())
    s = np.median(np.array(stage), axis=0)
    img = ctx.rt_download(sh, capi.PLANE_DEPTH)
    tris = sd.n_triangles
    verts = sum(len(m["positions"]) for m in sd.meshes)
    idx = sum(len(m["indices"]) for m in sd.meshes)
    b = size * size * 4 + 32 * verts + 4 * idx + 96 * len(sd.items)
    try:
        hbm = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        hbm = 6650.0
    gpu_ms = float(s[5])
    print(json.dumps({"workload": f"C3 shadow pass: {size}^2 shadow map, {tris} triangles, {len(sd.items)} draws", "gpu_ms_front_begin_to_tile_end": gpu_ms,
                      "vertex_setup_ms": float(s[0]), "binning_ms": float(s[1]), "tile_ms": float(s[2]), "wall_ms_median": statistics.median(wall),
                      "texels_written_below_1": int(np.count_nonzero(img < 1.0)), "algorithmic_bytes": b, "hbm_frac": b / 1e9 / (gpu_ms / 1e3) / hbm,
                      "mtri_per_s": tris / gpu_ms / 1e3, "iters": iters}))
    ctx.close()


if __name__ == "__main__":
    main()
