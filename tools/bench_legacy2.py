"""The legacy render-target demos at their shipped sizes (800x600 canvas, 2048^2 shadow map,
160x160 job tiles): one frame = shadow-map clear + shadow pass over every object + canvas / z-buffer (/ velocity) clear + lit pass
over every object (PCSS soft shadows, or Cook-Torrance + IBL + motion vectors), through shsb_legacy2_* / shsb_legacy3_*, timed on
the device with CUDA events on the context's stream, next to the reference's own draw_triangle_tile_* loops compiled from its
sources (oracle/_ref/libshs_legacy2_ref.so / libshs_legacy3_ref.so; the serial tile loop, one host thread) and checked against
them (shadow map / z-buffer / velocity bit-equal, canvas max LSB).
Usage (GPU box): python tools/bench_legacy2.py [frames] ; prints one JSON line per demo."""
import json
import os
import sys
import time

import numpy as np
from cuda import cudart  # CUDA events without importing torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import fuzz_cases  # noqa: E402
import test_legacy2_cpu as t2  # noqa: E402
import test_legacy3_cpu as t3  # noqa: E402
from test_zz_gpu_legacy2 import fill_uniforms  # noqa: E402
from leisure_software_renderer_b200 import capi, scenes  # noqa: E402
from leisure_software_renderer_b200.renderer import Context  # noqa: E402
from oracle.bindings import Legacy2Oracle, Legacy3Oracle  # noqa: E402  (the CPU baseline / checker leg only)

FLT_MAX = np.float32(np.finfo(np.float32).max)


def shipped_scene():
    """Floor grid + two Suzannes under the demos' light, at the demos' sizes; matrices are inputs of the path."""
    sc = fuzz_cases.legacy3_scene(1)
    sc.update(W=800, H=600, sm=2048, tile=(160, 160))
    m = scenes.load_suzanne()
    pos, nrm, uv = m["positions"][m["indices"]], m["normals"][m["indices"]], m["uvs"][m["indices"]]
    g = scenes.make_grid_plane(24.0, 8)
    objs = [(g["positions"][g["indices"]], g["normals"][g["indices"]], g["uvs"][g["indices"]] / 8.0, np.eye(4), (150, 160, 150, 255), True)]
    for k, (x, z, s) in enumerate(((-1.6, 0.5, 1.3), (1.5, 1.5, 1.0))):
        model = np.eye(4)
        model[:3, :3] *= s
        model[:3, 3] = (x, 1.2 * s, z)
        objs.append((pos, nrm, uv, model, (210 - 90 * k, 120, 60 + 120 * k, 255), False))
    sc["objs"] = objs
    sc["pbr"] = [(0.0, 0.7, 1.0), (0.8, 0.3, 1.0), (0.1, 0.5, 0.9)]
    rng = np.random.default_rng(5)
    sc["irradiance"] = rng.uniform(0.1, 1.0, (6, 16, 16, 3)).astype(np.float32)
    sc["prefiltered"] = [rng.uniform(0.1, 2.0, (6, 256 >> k, 256 >> k, 3)).astype(np.float32) for k in range(6)]
    sc["ibl_k"] = (0.30, 0.35, 1.0)
    return sc


def main():
    frames = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    gpu = Context(0)
    stream = gpu.stream()
    sc = shipped_scene()
    W, H, sm = sc["W"], sc["H"], sc["sm"]
    n_tris = sum(len(o[0]) // 3 for o in sc["objs"])
    meshes = [gpu.mesh_upload(pos, nrm, uv, None) for pos, nrm, uv, *_ in sc["objs"]]
    tex = gpu.texture_upload(sc["texture"])
    ibl = gpu.legacy3_ibl_upload(sc["irradiance"], sc["prefiltered"])
    clear = np.array([20, 20, 25, 255], np.uint8).view(np.uint32)
    zero2 = np.zeros(2, np.float32)
    for pbr in (False, True):
        sets = [(gpu.rt_create(capi.RT_COLOR_LDR, W, H), gpu.rt_create(capi.RT_DEPTH_MOTION, W, H), gpu.rt_create(capi.RT_SHADOW, sm, sm)) for _ in range(3)]
        us = []
        for k, (pos, nrm, uv, model, color, use_tex) in enumerate(sc["objs"]):
            u = fill_uniforms(sc, model, color, use_tex, tex, prev=pbr)
            if pbr:
                u.metallic, u.roughness, u.ao = sc["pbr"][k]
                u.ibl_diffuse_intensity, u.ibl_specular_intensity, u.ibl_reflection_strength = sc["ibl_k"]
            us.append(u)

        def frame(i):
            c, z, s = sets[i % len(sets)]
            gpu.rt_clear(s, capi.PLANE_DEPTH, FLT_MAX)
            for m, o in zip(meshes, sc["objs"]):
                gpu.legacy2_shadow_draw(m, sc["f32"](o[3]), sc["light_vp"], s, 160, 160)
            gpu.rt_clear(c, capi.PLANE_COLOR, clear)
            gpu.rt_clear(z, capi.PLANE_DEPTH, FLT_MAX)
            if pbr:
                gpu.rt_upload(z, capi.PLANE_MOTION, vel0)
            for m, u in zip(meshes, us):
                if pbr:
                    gpu.legacy3_draw_pbr(m, u, s, ibl, c, z)
                else:
                    gpu.legacy2_draw_softshadow(m, u, s, c, z)

        vel0 = np.zeros((H, W, 2), np.float32)
        for i in range(4):
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
        c, z, s = sets[(frames - 1) % len(sets)]
        g = [gpu.rt_download(s, capi.PLANE_DEPTH), gpu.rt_download(c), gpu.rt_download(z, capi.PLANE_DEPTH)] + ([gpu.rt_download(z, capi.PLANE_MOTION)] if pbr else [])
        t0 = time.perf_counter()
        r = (t3.render(Legacy3Oracle("reference"), sc) if pbr else t2.render(Legacy2Oracle("reference"), sc))
        ms_cpu = (time.perf_counter() - t0) * 1e3
        d = np.abs(g[1].astype(np.int32) - r[1].astype(np.int32)).max(axis=2)
        line = {"workload": f"legacy {'PBR / IBL' if pbr else 'soft-shadow'} demo frame: {n_tris} tris, {W}x{H}, {sm}^2 shadow map, 160x160 job tiles",
                "gpu_ms_per_frame": ms_total / frames, "gpu_frames_per_s": 1000.0 * frames / ms_total, "gpu_launches_per_frame": 4 * len(meshes) + 3 + (0 if not pbr else 0),
                "cpu_reference_ms_per_frame": ms_cpu, "cpu_threads": 1, "cpu_kind": "reference (draw_triangle_tile_* loops compiled from the demo source, serial over the job tiles)",
                "shadow_map_equal": bool(np.array_equal(g[0].view(np.uint32), r[0].view(np.uint32))), "zbuffer_equal": bool(np.array_equal(g[2].view(np.uint32), r[2].view(np.uint32))),
                "canvas_max_lsb_gpu_vs_reference": int(d.max()), "canvas_px_beyond_1_lsb": int(np.count_nonzero(d > 1)), "px_covered": int((r[2] < FLT_MAX).sum()), "frames_timed": frames}
        if pbr:
            line["velocity_equal"] = bool(np.array_equal(g[3].view(np.uint32), r[3].view(np.uint32)))
        print(json.dumps(line), flush=True)
        for rts in sets:
            for rt in rts:
                gpu.rt_destroy(rt)


if __name__ == "__main__":
    main()
