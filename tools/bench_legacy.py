"""BASELINE configs[0] "as shipped" (the legacy tile-job demo, SURVEY.md section 8a row L1 / 8d baseline B2): one frame = canvas +
z-buffer clear + one Suzanne through shsb_legacy_draw_blinn_phong at 640x480 (the BASELINE size) and at the demo's own 1240x980,
timed on the device with CUDA events on the context's stream, next to the reference's own RendererSystem::process -- one job per
80x80 tile on its ThreadedPriorityJobSystem -- compiled from the reference's sources (oracle/_ref/libshs_legacy_ref.so) and timed
on this box's host cores.  Usage (GPU box): python tools/bench_legacy.py [frames] ; prints one JSON line per size."""
import ctypes as C
import json
import os
import sys

import numpy as np
from cuda import cudart  # CUDA events without importing torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from leisure_software_renderer_b200 import capi, renderer, scenes  # noqa: E402
from leisure_software_renderer_b200.renderer import Context  # noqa: E402
from oracle.bindings import LegacyOracle  # noqa: E402  (the CPU baseline leg only)


def main():
    frames = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    gpu = Context(0)
    stream = gpu.stream()
    m = scenes.load_suzanne()
    mesh = gpu.mesh_upload(m["positions"], m["normals"], m["uvs"], m["indices"])
    pos, nrm = np.ascontiguousarray(m["positions"][m["indices"]]), np.ascontiguousarray(m["normals"][m["indices"]])
    cam, light = (0.0, 5.0, -20.0), tuple(float(v) for v in np.array([-1.0, -0.4, 1.0], np.float32) / np.float32(np.sqrt(2.16)))
    view, proj = renderer.legacy_camera(cam)
    model = renderer.legacy_world_matrix((0.0, 0.0, 10.0), (4.0, 4.0, 4.0), 0.0)
    mvp = renderer.legacy_mvp(proj, view, model)
    u = capi.LegacyUniforms(mvp, model, light, cam, (60, 100, 200, 255))
    ref = LegacyOracle("reference")
    cores = os.cpu_count() or 1
    for (w, h) in ((640, 480), (1240, 980)):
        sets = [(gpu.rt_create(capi.RT_COLOR_LDR, w, h), gpu.rt_create(capi.RT_SHADOW, w, h)) for _ in range(4)]
        black, fmax = np.array([0, 0, 0, 255], np.uint8).view(np.uint32), np.float32(np.finfo(np.float32).max)

        def frame(i):
            c, z = sets[i % len(sets)]
            gpu.rt_clear(c, capi.PLANE_COLOR, black)
            gpu.rt_clear(z, capi.PLANE_DEPTH, fmax)
            gpu.legacy_draw_blinn_phong(mesh, u, c, z)

        for i in range(8):
            frame(i)
        gpu.sync()
        _, e0 = cudart.cudaEventCreate()
        _, e1 = cudart.cudaEventCreate()
        cudart.cudaEventRecord(e0, stream)
        for i in range(frames):
            frame(i)
        cudart.cudaEventRecord(e1, stream)
        gpu.sync()
        _, ms_total = cudart.cudaEventElapsedTime(e0, e1)
        ms_gpu = ms_total / frames
        canvas_gpu = gpu.rt_download(sets[(frames - 1) % len(sets)][0])
        # CPU: the reference's own job system, all cores (the demo ships THREAD_COUNT = 20)
        canvas = np.zeros((h, w, 4), np.uint8)
        z = np.zeros((h, w), np.float32)
        f = ref.lib.shsref_legacy_frames_threaded
        f.restype = C.c_double
        cpu_frames = 20 if w <= 640 else 8
        ld, cp, col = np.ascontiguousarray(light, np.float32), np.ascontiguousarray(cam, np.float32), np.array([60, 100, 200, 255], np.uint8)
        ms_cpu = f(capi.fptr(pos), capi.fptr(nrm), C.c_uint32(len(pos)), capi.fptr(mvp), capi.fptr(model), capi.fptr(ld), capi.fptr(cp),
                   col.ctypes.data_as(C.POINTER(C.c_uint8)), w, h, 80, 80, cores, cpu_frames, canvas.ctypes.data_as(C.POINTER(C.c_uint8)), capi.fptr(z))
        d = np.abs(canvas_gpu.astype(np.int32) - canvas.astype(np.int32))
        print(json.dumps({"workload": f"C1 as shipped: Suzanne (967 tris) legacy Blinn-Phong tile-job frame {w}x{h}", "gpu_ms_per_frame": ms_gpu,
                          "gpu_frames_per_s": 1000.0 / ms_gpu, "gpu_launches_per_frame": 4, "cpu_reference_ms_per_frame": ms_cpu,
                          "cpu_reference_frames_per_s": 1000.0 / ms_cpu, "cpu_threads": cores, "cpu_kind": "reference (ThreadedPriorityJobSystem, one job per 80x80 tile)",
                          "canvas_max_lsb_gpu_vs_reference": int(d.max()), "frames_timed": frames}))
        for c, zz in sets:
            gpu.rt_destroy(c)
            gpu.rt_destroy(zz)


if __name__ == "__main__":
    main()
