"""The scene-level culling calls (csrc/scene_cull.cu) at a C2-like scale and beyond: N objects
(AABBs scattered around the camera), 1024 lights, 1080p / 16-px tiles.  Each call is timed end to end through the C-ABI with HOST
buffers (they are synchronous upload + kernels + download calls), wall clock over `reps` repetitions after warm-up, next to the
reference's own functions compiled from its sources (oracle/_ref/libshs_lightcull_ref.so, one host thread) and checked against them.
Usage (GPU box): python tools/bench_scene_cull.py [reps] ; prints one JSON line per (call, N)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from leisure_software_renderer_b200 import capi, scenes  # noqa: E402
from leisure_software_renderer_b200.renderer import Context  # noqa: E402
from oracle.bindings import LightCullReference, SceneCull  # noqa: E402  (the CPU baseline / checker leg only)


def timed(f, reps):
    f()
    t0 = time.perf_counter()
    for _ in range(reps):
        r = f()
    return (time.perf_counter() - t0) * 1e3 / reps, r


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    gpu = Context(0)
    ref, lref = SceneCull("reference"), LightCullReference()
    rng = np.random.default_rng(1)
    w, h, ts, zn, zf = 1920, 1080, 16, 0.1, 200.0
    eye, tgt = (0.0, 6.0, -30.0), (0.0, 1.0, 0.0)
    vp = scenes.camera_viewproj(eye, tgt, (0.0, 1.0, 0.0), float(np.radians(60.0)), w / h, zn, zf)
    view = np.ascontiguousarray(scenes.look_at_lh(eye, tgt, (0.0, 1.0, 0.0)).astype(np.float32).T).reshape(16)
    lights = scenes.make_lights(768, 256, (-40.0, 0.0, -40.0), (40.0, 8.0, 40.0), seed=3, range_lo=2.0, range_hi=9.0)
    all_lights = np.arange(len(lights), dtype=np.uint32)
    gpu.lights_upload(lights)
    desc = capi.LightCullDesc(vp, w, h, capi.LIGHT_CULL_CLUSTERED, ts, 256, depth_slices=16, z_near=zn, z_far=zf)
    gpu.light_cull_ex(desc)
    for n in (1000, 10000, 100000):
        c = rng.uniform(-60, 60, (n, 3)) * np.array([1, 0.15, 1])
        half = np.abs(rng.normal(0, 0.8, (n, 3))) + 0.05
        aabbs = np.concatenate([c - half, c + half], axis=1).astype(np.float32)
        bounds = lref.bounds(aabbs)
        vis_objects = np.arange(n, dtype=np.uint32)
        cpu_n = min(n, 10000)     # the reference's serial functions on a bounded sample, scaled
        calls = {
            "cull_objects_frustum": (lambda: gpu.cull_objects_frustum(bounds, vp), lambda: ref.cull_objects(aabbs[:cpu_n], vp)),
            "collect_object_lights(SphereAabb, 1024 lights)": (lambda: gpu.collect_object_lights(aabbs, all_lights, lights, 1),
                                                               lambda: ref.collect_object_lights(aabbs[:cpu_n], all_lights, lights, 1)),
            "select_object_lights_from_bins(clustered 120x68x16)": (lambda: gpu.select_object_lights_from_bins(aabbs, view, vp, True, zn, zf, lights, 1), None),
            "tile_depth_range_from_scene(1080p, 16 px)": (lambda: gpu.tile_depth_range_from_scene(aabbs, vis_objects, view, vp, w, h, ts, zn, zf),
                                                          lambda: ref.tile_depth_range_from_scene(aabbs[:cpu_n], vis_objects[:cpu_n], view, vp, w, h, ts, zn, zf)),
        }
        for name, (g, c_) in calls.items():
            ms_gpu, rg = timed(g, reps)
            line = {"call": name, "objects": n, "gpu_ms_per_call_host_buffers": ms_gpu, "objects_per_s": n / ms_gpu * 1e3, "reps": reps}
            if c_ is not None:
                ms_cpu, rc = timed(c_, 2)
                line.update({"cpu_reference_ms_scaled_to_n": ms_cpu * n / cpu_n, "cpu_sample_objects": cpu_n, "cpu_threads": 1,
                             "cpu_kind": "reference (compiled from its headers, serial)"})
                if cpu_n == n:
                    line["equal_to_reference"] = bool(all(np.array_equal(np.asarray(a).view(np.uint8), np.asarray(b).view(np.uint8)) for a, b in zip(rg, rc)))
            print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
