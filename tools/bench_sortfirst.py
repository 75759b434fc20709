"""Strong scaling of ONE 8K frame (BASELINE configs[4]: 7680x4320, ~10 M triangles, 1024 lights) split sort-first over the
screen across N GPUs of one box:  every rank holds the whole scene, renders only the 16-px tile rows
it owns (ShsbFrameParams::own_row_*) and rank 0 assembles the LDR frame over NVLink (NCCL send / recv of disjoint row ranges,
on a side stream so the assembly of frame i overlaps the rendering of frame i + 1).

    python tools/bench_sortfirst.py                                   # N = 1: the whole frame on one GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_sortfirst.py [--layout bands|interleaved]

bands       : contiguous bands of tile rows, boundaries chosen so that every band carries the same share of a per-row cost
              measured on a calibration frame (shaded pixels per tile row + a constant per row), which is what lets the
              library skip the draws that cannot reach a band (most of the scene for each rank);
interleaved : rank r owns stripes of 2 tile rows every 2 N rows (balanced by construction, but every rank sets up every triangle).
One JSON line from rank 0: frames/s (CUDA events, max over ranks), per-rank tile rows and shaded fragments.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

from leisure_software_renderer_b200 import capi, scenes
from leisure_software_renderer_b200.renderer import Context


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layout", default="bands", choices=["bands", "interleaved"])
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--config", default="c5", choices=["c5", "c2"])
    ap.add_argument("--emulate", type=int, nargs=2, metavar=("WORLD", "RANK"), default=None,
                    help="single process: render only the partition rank RANK of WORLD would own (no frame assembly) -- the period one rank can sustain")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = Context(local)
    sd = scenes.scene_c5() if args.config == "c5" else scenes.scene_c2()
    W, H = sd.w, sd.h
    for m in sd.meshes:
        ctx.mesh_upload(m["positions"], m["normals"], m["uvs"], m["indices"])
    ctx.lights_upload(sd.lights.view(np.uint8))
    NSETS = 3
    sets = [(ctx.rt_create(capi.RT_COLOR_HDR, W, H), ctx.rt_create(capi.RT_DEPTH_MOTION, W, H, sd.zn, sd.zf), ctx.rt_create(capi.RT_COLOR_LDR, W, H)) for _ in range(NSETS)]
    tile_rows = (H + 15) // 16

    # ---- calibration frame (whole frame, every rank computes the same thing): cost per tile row
    whole = ctx.frame_forward_plus(sd.scene, sd.fp, *sets[0]).as_dict()
    depth = ctx.rt_download(sets[0][1], capi.PLANE_DEPTH)
    shaded_rows = (depth[::-1] < 1.0).sum(axis=1)                                   # per pixel row, from the top
    pad = tile_rows * 16 - H
    px_cost = np.concatenate([shaded_rows, np.zeros(pad, shaded_rows.dtype)]).reshape(tile_rows, 16).sum(axis=1).astype(np.float64)
    del depth
    # triangles per tile row: every draw's triangle count at the tile row its origin projects to (instances are small on screen)
    vp = np.asarray(sd.viewproj, dtype=np.float64).reshape(4, 4).T                     # row-major
    tri_cost = np.zeros(tile_rows)
    for it in sd.items:
        m = sd.meshes[it["mesh"] - 1]
        c = vp @ np.array([*it["pos"], 1.0])
        if c[3] <= 1e-3:
            continue
        py = (c[1] / c[3] * 0.5 + 0.5) * (H - 1)
        ty = int(np.clip((H - 1 - py) // 16, 0, tile_rows - 1))
        tri_cost[ty] += len(m["indices"]) // 3
    # a frame's time is about half per-pixel work (shading) and half per-triangle work (set-up, binning, coverage)
    cost = px_cost / max(px_cost.sum(), 1.0) + tri_cost / max(tri_cost.sum(), 1.0) + 0.02 / tile_rows
    fp = capi.FrameParams.from_buffer_copy(sd.fp)
    emu_world, emu_rank = (args.emulate if args.emulate else (world, rank))
    if emu_world > 1:
        part_world, part_rank = emu_world, emu_rank
        if args.layout == "bands":
            cum = np.concatenate([[0.0], np.cumsum(cost)])
            cuts = [int(np.searchsorted(cum, cum[-1] * r / part_world)) for r in range(part_world + 1)]
            cuts[0], cuts[-1] = 0, tile_rows
            for r in range(1, part_world + 1):
                cuts[r] = max(cuts[r], cuts[r - 1] + 1) if r < part_world else tile_rows
            first, count, stride = cuts[part_rank], cuts[part_rank + 1] - cuts[part_rank], 1 << 24
        else:
            first, count, stride = part_rank * 2, 2, part_world * 2
        fp.own_row_first, fp.own_row_count, fp.own_row_stride = first, count, stride
    else:
        first, count, stride = 0, tile_rows, 1 << 24
    owned = np.array([ty >= first and (ty - first) % stride < count for ty in range(tile_rows)])

    # pixel rows (framebuffer order, bottom-origin) this rank owns, as a torch index for packing the LDR rows
    top_tile_of_fb_row = (H - 1 - np.arange(H)) // 16
    my_rows = torch.from_numpy(np.nonzero(owned[top_tile_of_fb_row])[0]).to(f"cuda:{local}")
    n_rows = torch.tensor([my_rows.numel()], device=f"cuda:{local}")
    rows_of = [torch.zeros_like(n_rows) for _ in range(world)]
    if world > 1:
        dist.all_gather(rows_of, n_rows)
    rows_of = [int(t.item()) for t in rows_of] if world > 1 else [H]
    all_rows = None
    if world > 1 and rank == 0:
        all_rows = []
        for r in range(world):
            if args.layout == "bands":
                o = np.array([cuts[r] <= ty < cuts[r + 1] for ty in range(tile_rows)])
            else:
                o = np.array([ty >= r * 2 and (ty - r * 2) % (world * 2) < 2 for ty in range(tile_rows)])
            all_rows.append(torch.from_numpy(np.nonzero(o[top_tile_of_fb_row])[0]).to("cuda:0"))

    def ldr_view(rt):
        ptr, nbytes = ctx.rt_device_ptr(rt, capi.PLANE_COLOR)

        class _Cai:
            __cuda_array_interface__ = {"shape": (H, W * 4), "typestr": "|u1", "data": (ptr, False), "version": 3}
        return torch.as_tensor(_Cai(), device=torch.device("cuda", local))

    views = [ldr_view(s[2]) for s in sets]
    stream = torch.cuda.ExternalStream(ctx.stream(), device=local)
    comm = torch.cuda.Stream(device=local)
    frame_out = torch.empty((H, W * 4), dtype=torch.uint8, device=f"cuda:{local}") if rank == 0 else None
    recv_bufs = [torch.empty((rows_of[r], W * 4), dtype=torch.uint8, device="cuda:0") for r in range(world)] if (rank == 0 and world > 1) else None
    gather_done = [None] * NSETS
    comm_events = []
    pack_buf = torch.empty((my_rows.numel(), W * 4), dtype=torch.uint8, device=f"cuda:{local}") if (world > 1 and rank != 0 and args.layout != "bands") else None
    # framebuffer row range [a, b) of every rank's band (rows are bottom-origin, tile rows count from the top)
    fb_range = None
    if world > 1 and args.layout == "bands":
        fb_range = [(max(0, H - cuts[r + 1] * 16), H - cuts[r] * 16) for r in range(world)]

    def step(i):
        k = i % NSETS
        if gather_done[k] is not None:
            stream.wait_event(gather_done[k])
        ctx.frame_forward_plus(sd.scene, fp, *sets[k], want_stats=False)
        if world > 1:
            ev = torch.cuda.Event(); ev.record(stream)
            with torch.cuda.stream(comm):
                comm.wait_event(ev)
                c0 = torch.cuda.Event(enable_timing=True); c0.record(comm)
                if args.layout == "bands":
                    # a band is a contiguous range of framebuffer rows: receive straight into the assembled frame, all peers in ONE grouped NCCL call
                    if rank == 0:
                        a, b = fb_range[0]
                        frame_out[a:b].copy_(views[k][a:b], non_blocking=True)
                        ops = [dist.P2POp(dist.irecv, frame_out[fb_range[r][0]:fb_range[r][1]], r) for r in range(1, world)]
                        for q in dist.batch_isend_irecv(ops):
                            q.wait()
                    else:
                        a, b = fb_range[rank]
                        for q in dist.batch_isend_irecv([dist.P2POp(dist.isend, views[k][a:b], 0)]):
                            q.wait()
                elif rank == 0:
                    frame_out[all_rows[0]] = views[k][all_rows[0]]
                    ops = [dist.P2POp(dist.irecv, recv_bufs[r], r) for r in range(1, world)]
                    for q in dist.batch_isend_irecv(ops):
                        q.wait()
                    for r in range(1, world):
                        frame_out[all_rows[r]] = recv_bufs[r]
                else:
                    pack_buf.copy_(views[k][my_rows])
                    for q in dist.batch_isend_irecv([dist.P2POp(dist.isend, pack_buf, 0)]):
                        q.wait()
                done = torch.cuda.Event(enable_timing=True); done.record(comm)
                comm_events.append((c0, done))
            gather_done[k] = done

    def barrier():
        torch.cuda.synchronize(); ctx.sync()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    st = ctx.frame_forward_plus(sd.scene, fp, *sets[0]).as_dict()
    sm = ctx.last_stage_ms()  # stage times of this synchronous frame (this rank's partition alone on its GPU)
    for i in range(args.warmup):
        step(i)
    barrier()
    import time
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.host_submit_us(reset=True)
    e0.record(stream)
    t_host = time.perf_counter()
    for i in range(args.steps):
        step(i)
    host_ms = (time.perf_counter() - t_host) * 1e3 / args.steps
    hu = ctx.host_submit_us(reset=True)
    lib_ms = float(hu[:5].sum() / max(hu[5], 1.0)) / 1e3
    for d in gather_done:
        if d is not None:
            stream.wait_event(d)
    e1.record(stream)
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    comm_ms = float(np.mean([a.elapsed_time(b) for a, b in comm_events[-args.steps:]])) if comm_events else 0.0
    mine = torch.tensor([float(owned.sum()), float(st["frag_shaded"]), float(st["tri_input"]), host_ms, lib_ms, float(sm[0]), float(sm[1]), float(sm[2]), comm_ms],
                        dtype=torch.float64, device="cuda")
    per_rank = [torch.zeros_like(mine) for _ in range(world)]
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_gather(per_rank, mine)
    else:
        per_rank = [mine]
    ok = None
    if rank == 0 and world > 1:
        # the assembled frame against this rank's own whole-frame rendering of the same scene
        ctx.frame_forward_plus(sd.scene, sd.fp, *sets[1])
        ok = bool(torch.equal(frame_out, views[1]))
    if rank == 0:
        t = float(ms.item()) / args.steps
        print(json.dumps({"metric": "frames/s", "config": sd.name, "resolution": [W, H], "n_gpus": world, "layout": args.layout if emu_world > 1 else "whole frame", "emulated_partition": args.emulate,
                          "host_threads": os.environ.get("SHSB_HOST_THREADS", "default: clamp(cores / 8, 1, 4)"),
                          "value": 1e3 / t, "ms_per_frame": t, "steps": args.steps, "scaling": "strong",
                          "tile_rows_per_rank": [int(p[0].item()) for p in per_rank], "frag_shaded_per_rank": [int(p[1].item()) for p in per_rank],
                          "tri_input_per_rank": [int(p[2].item()) for p in per_rank],
                          "host_submit_ms_per_rank": [round(float(p[3].item()), 3) for p in per_rank], "library_host_ms_per_rank": [round(float(p[4].item()), 3) for p in per_rank],
                          "gpu_stage_ms_per_rank": {"vertex_clip_setup": [round(float(p[5].item()), 3) for p in per_rank], "binning": [round(float(p[6].item()), 3) for p in per_rank],
                                                    "tile": [round(float(p[7].item()), 3) for p in per_rank]},
                          "frame_assembly_ms_per_rank": [round(float(p[8].item()), 3) for p in per_rank],
                          "whole_frame_stats": whole,
                          "assembled_frame_equals_whole_frame": ok}), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
