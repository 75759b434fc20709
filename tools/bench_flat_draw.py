"""shsb_flat_draw_blinn_phong / shsb_flat_draw_multi_light next to the reference's own draws (oracle/_ref/libshs_flat_draw_ref.so:
debug_draw::draw_mesh_blinn_phong_transformed and the demo's draw_mesh_multi_light_transformed, one host thread like the demo's frame
loop): python tools/bench_flat_draw.py [reps] -> one JSON line per scene size and draw kind.  Scene: the demo's 1200 x 900 canvas
(exp-plumbing/hello_light_types_culling_sw.cpp:43-46), a floor under n tessellated spheres / boxes, 64 lights of the four models, 8
selected per object."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np

import fuzz_cases
from leisure_software_renderer_b200 import capi, scenes
from leisure_software_renderer_b200.renderer import Context
from oracle import bindings


def scene(n, seed=0, w=1200, h=900):
    rng = np.random.default_rng(seed)
    box_v = np.array([[x, y, z] for z in (-.5, .5) for y in (-.5, .5) for x in (-.5, .5)], np.float32)
    box_i = np.array([0, 1, 3, 0, 3, 2, 4, 6, 7, 4, 7, 5, 0, 4, 5, 0, 5, 1, 2, 3, 7, 2, 7, 6, 0, 2, 6, 0, 6, 4, 1, 5, 7, 1, 7, 3], np.uint32)
    quad_v = np.array([[-.5, 0, -.5], [.5, 0, -.5], [.5, 0, .5], [-.5, 0, .5]], np.float32)
    quad_i = np.array([0, 2, 1, 0, 3, 2], np.uint32)
    sph_v, sph_i = fuzz_cases._uv_sphere(24, 16)
    meshes = [(quad_v, quad_i), (box_v, box_i), (sph_v, sph_i)]
    table, fi, bv = [], 0, 0
    for v, i in meshes:
        table.append([fi, len(i), bv]); fi += len(i); bv += len(v)
    draw_mesh, models, base = [0], [np.diag([96.0, 1.0, 96.0, 1.0]).astype(np.float32).T.reshape(16)], [[0.5, 0.5, 0.55]]
    side = max(1, int(np.ceil(np.sqrt(n))))
    for k in range(n):
        M = np.eye(4)
        M[:3, :3] = fuzz_cases._rotation(rng) @ np.diag(rng.uniform(0.6, 1.6, 3))
        M[:3, 3] = [(k % side - side / 2) * 44.0 / side, rng.uniform(0.8, 3.0), (k // side - side / 2) * 44.0 / side]
        draw_mesh.append(1 + int(rng.integers(0, 2))); models.append(M.T.astype(np.float32).reshape(16)); base.append(rng.uniform(0.2, 1.0, 3))
    lights = fuzz_cases.flat_draw_lights(rng, 64, 22.0)
    eye, tgt = (0.0, 14.0, -30.0), (0.0, 0.0, 0.0)
    nd = len(draw_mesh)
    canvas = np.zeros((h, w, 4), np.uint8); canvas[..., 3] = 255
    return {"draw_mesh": np.array(draw_mesh, np.uint32), "models": np.array(models, np.float32), "base": np.array(base, np.float32), "sel_counts": np.full(nd, 8, np.uint32),
            "sel_idx": rng.integers(0, 64, (nd, 8)).astype(np.uint32), "mesh_table": np.array(table, np.uint32), "vertices": np.concatenate([m[0] for m in meshes]),
            "indices": np.concatenate([m[1] for m in meshes]), "view_proj": scenes.camera_viewproj(eye, tgt, (0, 1, 0), float(np.radians(60)), w / h, 0.05, 300.0),
            "camera": np.array(eye, np.float32), "light_dir": np.array([0.3, -1.0, 0.2], np.float32), "lights": lights, "W": w, "H": h, "canvas": canvas, "depth": np.ones((h, w), np.float32)}


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    ctx = Context(0)
    ref = bindings.FlatDraw("reference") if bindings.FlatDraw.available() else None
    for n in (100, 1000, 10000):
        sc = scene(n)
        meshes = [ctx.mesh_upload(sc["vertices"][int(b):], indices=sc["indices"][int(f):int(f) + int(c)]) for f, c, b in sc["mesh_table"]]
        canvas, depth = ctx.rt_create(capi.RT_COLOR_LDR, sc["W"], sc["H"]), ctx.rt_create(capi.RT_SHADOW, sc["W"], sc["H"])
        draws = ctx._flat_draws([{"mesh": meshes[int(m)], "model": sc["models"][i], "base_color": sc["base"][i], "selection": sc["sel_idx"][i]} for i, m in enumerate(sc["draw_mesh"])])
        vp, cam, ld = capi.fptr(sc["view_proj"]), capi.fptr(sc["camera"]), capi.fptr(sc["light_dir"])
        one = np.float32(1.0)
        tris = int(sum(int(sc["mesh_table"][int(m)][1]) // 3 for m in sc["draw_mesh"]))
        for mode, name in ((0, "blinn_phong"), (1, "multi_light (8 of 64 lights per object)")):
            def frame():
                ctx.rt_clear(depth, capi.PLANE_DEPTH, one)
                if mode == 0:
                    rc = ctx.lib.shsb_flat_draw_blinn_phong(ctx.h, draws, len(draws), vp, cam, ld, canvas, depth)
                else:
                    rc = ctx.lib.shsb_flat_draw_multi_light(ctx.h, draws, len(draws), vp, cam, sc["lights"].ctypes.data_as(capi.C.c_void_p), len(sc["lights"]), canvas, depth)
                assert rc == 0, rc
            ctx.rt_upload(canvas, capi.PLANE_COLOR, sc["canvas"])
            frame(); frame()
            ctx.sync()
            t0 = time.perf_counter()
            for _ in range(reps):
                frame()
            ctx.sync()
            gpu_ms = (time.perf_counter() - t0) * 1e3 / reps
            got_c, got_d = ctx.rt_download(canvas).reshape(sc["H"], sc["W"], 4), ctx.rt_download(depth, capi.PLANE_DEPTH).reshape(sc["H"], sc["W"])
            line = {"call": f"flat_draw_{name}, {sc['W']}x{sc['H']}", "objects": n, "triangles": tris, "covered_texels": int(np.count_nonzero(got_d < 1.0)),
                    "gpu_ms_per_frame_incl_depth_clear": gpu_ms, "reps": reps}
            if ref is not None:
                t0 = time.perf_counter()
                want_c, want_d = ref.run(sc, mode)
                line["cpu_reference_ms"] = (time.perf_counter() - t0) * 1e3
                line["cpu_kind"] = "reference (its own text compiled, 1 thread, like the demo's frame loop; includes copying the 1200x900 targets in and out)"
                line["depth_bit_equal"] = bool(np.array_equal(got_d.view(np.uint32), want_d.view(np.uint32)))
                line["colour_channels_off_by_1"] = int(np.count_nonzero(np.abs(got_c.astype(np.int16) - want_c.astype(np.int16)) == 1))
                line["colour_max_diff"] = int(np.abs(got_c.astype(np.int16) - want_c.astype(np.int16)).max())
            print(json.dumps(line), flush=True)
        ctx.rt_destroy(canvas); ctx.rt_destroy(depth)
        for m in meshes:
            ctx.lib.shsb_mesh_destroy(ctx.h, m)
    ctx.close()


if __name__ == "__main__":
    main()
