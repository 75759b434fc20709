"""shsb_software_occlusion next to the reference's own run_software_occlusion_pass (oracle/_ref/libshs_occlusion_ref.so, one host
thread): python tools/bench_occlusion.py [reps] -> one JSON line per scene size.  Scene: a street of box occluders (walls near the
camera hide most of what is behind them), 320 x 180 occlusion buffer like exp-plumbing/hello_light_types_culling_sw.cpp."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np

from leisure_software_renderer_b200 import scenes
from leisure_software_renderer_b200.renderer import Context
from oracle import bindings


def scene(n, seed=0):
    rng = np.random.default_rng(seed)
    box_v = np.array([[x, y, z] for z in (-.5, .5) for y in (-.5, .5) for x in (-.5, .5)], np.float32)
    box_i = np.array([0, 1, 3, 0, 3, 2, 4, 6, 7, 4, 7, 5, 0, 4, 5, 0, 5, 1, 2, 3, 7, 2, 7, 6, 0, 2, 6, 0, 6, 4, 1, 5, 7, 1, 7, 3], np.uint32)
    aabbs, models = [], []
    for i in range(n):
        c = np.array([rng.uniform(-60, 60), rng.uniform(0, 4), rng.uniform(2, 200)])
        half = rng.uniform(0.3, 1.5, 3) * (6.0 if rng.random() < 0.08 else 1.0)
        m = np.eye(4, dtype=np.float32)
        m[0, 0], m[1, 1], m[2, 2] = 2 * half
        m[:3, 3] = c
        aabbs.append(np.concatenate([c - half, c + half])); models.append(m.T.reshape(16))
    eye, tgt = (0.0, 3.0, -5.0), (0.0, 2.0, 50.0)
    vp = scenes.camera_viewproj(eye, tgt, (0, 1, 0), float(np.radians(60)), 320 / 180, 0.1, 1000.0)
    view = np.ascontiguousarray(scenes.look_at_lh(eye, tgt, (0.0, 1.0, 0.0)).astype(np.float32).T).reshape(16)
    return {"aabbs": np.array(aabbs, np.float32), "visible": np.arange(n, dtype=np.uint32), "object_mesh": np.zeros(n, np.uint32), "models": np.array(models, np.float32),
            "mesh_table": np.array([[0, 36, 0]], np.uint32), "vertices": box_v, "indices": box_i, "view": view, "view_proj": vp, "occ_w": 320, "occ_h": 180, "eps": 1e-4}


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    ctx = Context(0)
    ref = bindings.SoftwareOcclusion("reference") if bindings.SoftwareOcclusion.available() else None
    for n in (100, 1000, 10000):
        sc = scene(n)
        args = (sc["aabbs"], sc["visible"], sc["object_mesh"], sc["models"], sc["mesh_table"], sc["vertices"], sc["indices"], sc["view"], sc["view_proj"], 320, 180, 1e-4)
        got = ctx.software_occlusion(*args)
        t0 = time.perf_counter()
        for _ in range(reps):
            ctx.software_occlusion(*args)
        gpu_ms = (time.perf_counter() - t0) * 1e3 / reps
        line = {"call": "software_occlusion(320x180, box occluders)", "objects": n, "visible": int(got[2][2]), "occluded": int(got[2][3]), "gpu_ms_per_call_host_buffers": gpu_ms, "reps": reps}
        if ref is not None:
            want = ref.run(sc)
            t0 = time.perf_counter()
            for _ in range(max(1, reps // 2)):
                ref.run(sc)
            line["cpu_reference_ms"] = (time.perf_counter() - t0) * 1e3 / max(1, reps // 2)
            line["cpu_kind"] = "reference (compiled from its header, 1 thread)"
            line["equal_to_reference"] = bool(np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]) and np.array_equal(got[3] + 0, want[3] + 0))
        print(json.dumps(line), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
