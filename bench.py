#!/usr/bin/env python
"""bench.py -- frames/s of the 1080p Forward+ frame (BASELINE.json configs[1]) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W]            this repo's CUDA path
    python bench.py --impl reference [--gpus N] --steps K ...      the reference's own CPU code (oracle/_ref)

A "step" is one pass of the hot path over one frame of the synthetic scene: light culling (tile
light-list build) -> vertex/clip/setup -> binning -> tile raster + Forward+ shading -> tonemap,
through shsb_frame_forward_plus (include/shsb.h).  With N > 1 (torchrun, one process per GPU) the work
is sort-first over a camera batch: rank r renders camera r of a ring around the same scene and the
LDR frames are gathered on rank 0 over NCCL -- per-GPU work is fixed, so scaling is "weak".

value : whole-job frames/s with all inputs resident in HBM (device events, max over ranks).
e2e   : the same metric through the C-ABI with HOST buffers -- per step the light records are uploaded
        from pinned host memory, the draw list goes H2D inside the call, and the LDR frame is read back
        to pinned host memory.
roofline : the tile raster+shade kernel's algorithmic bytes / its mean launch time (library-recorded CUDA
        events on the launching stream) against the measured HBM peak in MEASURED_PEAKS.json.
cpu_baseline : the CPU oracle (port) timed on this box for a bounded sample of the same frames.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

W, H = 1920, 1080
TIMING_STRIDE = int(os.environ.get("SHSB_BENCH_STRIDE", "8"))  # frames between two frames whose per-stage CUDA events are recorded inside the timed region
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
            t0 = time.time()
            while not self.lines and time.time() - t0 < 3.0:  # nvidia-smi needs a moment before its first sample
                time.sleep(0.01)
            self.lines.clear()                                   # keep only samples taken inside the timed region
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE tile_kernel launch, from the committed `ncu --set full`
    capture summarised under profiles/ (not measured live: a run under ncu is never a bench run)."""
    p = os.path.join(ROOT, "profiles", "tile_kernel_traffic.json")
    try:
        t = json.load(open(p))
        return int(t["dram_bytes_read"]) + int(t["dram_bytes_write"]), t.get("capture")
    except Exception:
        return None, None


def dist_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def algorithmic_bytes(sd, counts, n_unique_vertices, n_indices):
    """every output byte written once, every input byte read once."""
    px = sd.w * sd.h
    tiles = counts.size
    list_entries = int(np.minimum(counts, sd.fp.max_lights_per_tile).sum())
    n_lights = 0 if sd.lights is None else len(sd.lights)
    frame = px * 24 + 32 * n_unique_vertices + 4 * n_indices + 96 * len(sd.items) + 160 * n_lights + 2 * (4 * tiles + 4 * list_entries)
    tile_kernel = px * 24 + (4 * tiles + 4 * list_entries) + 160 * n_lights
    return frame, tile_kernel


def run_ours(args):
    rank, local_rank, world = dist_env()
    import torch
    import torch.distributed as dist
    from leisure_software_renderer_b200 import build, capi, scenes
    from leisure_software_renderer_b200.renderer import Context

    build.build()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the raster path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # NCCL chatter (e.g. its version line) must not share stdout with the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    ctx = Context(local_rank)
    base = scenes.scene_c2(W, H)
    sd = scenes.camera_ring(base, world)[rank] if world > 1 else base
    for m in sd.meshes:
        ctx.mesh_upload(m["positions"], m["normals"], m["uvs"], m["indices"])
    lights_pinned = torch.from_numpy(np.ascontiguousarray(sd.lights).view(np.uint8).copy()).pin_memory()
    ctx.lights_upload(lights_pinned.numpy())
    fp = sd.fp

    # Rotating render-target sets: 4 x (33.2 + 8.3 + 8.3 MB) = 199 MB > 126 MB L2, so a frame's output lines
    # cannot still be dirty-resident in L2 when the same buffers are written again.
    sets = [(ctx.rt_create(capi.RT_COLOR_HDR, W, H), ctx.rt_create(capi.RT_DEPTH_MOTION, W, H, sd.zn, sd.zf), ctx.rt_create(capi.RT_COLOR_LDR, W, H))
            for _ in range(NSETS)]
    stream = torch.cuda.ExternalStream(ctx.stream(), device=local_rank)

    def ldr_tensor(rt):
        ptr, nbytes = ctx.rt_device_ptr(rt, capi.PLANE_COLOR)

        class _Cai:
            __cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}
        return torch.as_tensor(_Cai(), device=torch.device("cuda", local_rank))

    ldr_views = [ldr_tensor(s[2]) for s in sets]
    gather_bufs = [torch.empty_like(ldr_views[0]) for _ in range(world)] if (world > 1 and rank == 0) else None
    host_ldr = [torch.empty(W * H * 4, dtype=torch.uint8).pin_memory() for _ in range(2)]

    comm_stream = torch.cuda.Stream(device=local_rank) if world > 1 else None
    gather_done = [None] * NSETS   # per render-target set: the gather that last read its LDR plane
    last_gather = [None]

    def step(i, e2e=False):
        k = i % NSETS
        hdr, dm, ldr = sets[k]
        if e2e:
            ctx.lights_upload(lights_pinned.numpy())          # H2D 160 B x n_lights from pinned memory
        if world > 1 and gather_done[k] is not None:
            stream.wait_event(gather_done[k])                 # do not overwrite an LDR plane that is still being gathered
            ctx.fence()
        ctx.frame_forward_plus(sd.scene, fp, hdr, dm, ldr, want_stats=False)   # asynchronous; draw list H2D inside the call
        if world > 1:
            # frame assembly over NVLink: the gather of frame i runs on its own stream and overlaps with frame i+1
            ctx.fence()
            ev = torch.cuda.Event()
            ev.record(stream)
            with torch.cuda.stream(comm_stream):
                comm_stream.wait_event(ev)
                dist.gather(ldr_views[k], gather_bufs, dst=0)
                done = torch.cuda.Event()
                done.record(comm_stream)
            gather_done[k] = done
            last_gather[0] = done
        if e2e:
            # D2H of the step's result into pinned memory (copy stream: overlaps with the next step's rendering)
            ctx.rt_download_async(ldr, capi.PLANE_COLOR, host_ldr[i % 2].data_ptr(), W * H * 4)

    def barrier():
        torch.cuda.synchronize()
        ctx.sync()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # one frame with statistics (also sizes the per-frame arena), then warm-up
    st = ctx.frame_forward_plus(sd.scene, fp, *sets[0]).as_dict()
    counts, _ = ctx.light_lists_download()
    n_warm = max(args.warmup, 3)

    def warm(e2e):
        """Warm-up of the path that is timed next: the same calls as the timed loop (both executable-graph topologies of the
        front end -- with and without stage events --, the light-upload ring, the read-back streams and their pinned buffers),
        at least 2 x NSETS frames so that every render-target set and every transient arena has been through it once."""
        if not e2e:
            ctx.timing_enable(TIMING_STRIDE)
        for i in range(max(n_warm, 2 * NSETS)):
            step(i, e2e)
        if e2e:
            ctx.sync()
        barrier()
        if not e2e:
            ctx.timing_collect()
            ctx.timing_enable(False)

    def timed(e2e):
        warm(e2e)
        launches0 = ctx.launch_count()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if not e2e:
            ctx.timing_enable(TIMING_STRIDE)   # stage events on every 8th frame only: they are extra commands on the critical stream
        ctx.host_submit_us(reset=True)
        e0.record(stream)
        t_host = time.perf_counter()
        for i in range(args.steps):
            step(i, e2e)
        host_ms = (time.perf_counter() - t_host) * 1e3 / args.steps   # host time to SUBMIT a step (includes back-pressure waits on the staging ring)
        hu = ctx.host_submit_us(reset=True)
        host_ms = {"total_ms": host_ms, "library_us": {k: float(hu[i] / max(hu[5], 1.0)) for i, k in
                   enumerate(("draw_list", "staging_incl_ring_wait", "arena_checks", "capture_enqueue", "graph_update_launch_tile_launch"))}}
        if world > 1 and last_gather[0] is not None:
            stream.wait_event(last_gather[0])  # the last frame assembly belongs to the timed region
        if e2e:
            ctx.sync()  # the last read-backs (copy stream) belong to the timed region
        ctx.fence()     # frames run on several render streams: the main stream (and e1) behind all of them
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        stages = ctx.timing_collect() if not e2e else None
        if not e2e:
            ctx.timing_enable(False)
        clocks = sampler.stop() if rank == 0 else None
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), stages, clocks, ctx.launch_count() - launches0, host_ms

    ms_dev, stages, clocks, launches, host_dev = timed(False)
    ms_e2e, _, clocks_e2e, _, host_e2e = timed(True)

    n_items = len(sd.items)
    n_blocks = sum((len(sd.meshes[it["mesh"] - 1]["indices"]) // 3 + 127) // 128 for it in sd.items)
    h2d = n_items * 192 + n_blocks * 8 + len(sd.lights) * 160
    d2h = W * H * 4
    fps = world * args.steps / (ms_dev / 1e3)
    fps_e2e = world * args.steps / (ms_e2e / 1e3)
    hbm, peak_src = peaks()
    mesh = sd.meshes[0]
    b_frame, b_tile = algorithmic_bytes(sd, counts, len(mesh["positions"]), len(mesh["indices"]))
    stage_mean = stages.mean(axis=0) if stages is not None and len(stages) else np.zeros(4)
    tile_ms = float(stage_mean[2])

    if rank == 0:
        line = {
            "metric": "frames/s", "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": workload_config(sd, world),
            "mtri_per_s": fps * st["tri_input"] / 1e6, "mfrag_per_s": fps * st["frag_covered"] / 1e6,
            "frame_stats": st,
            "e2e": {"value": fps_e2e, "unit": "frames/s", "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
                    "ms_per_step": ms_e2e / args.steps, "host_submit_ms_per_step": host_e2e},
            "host_submit_ms_per_step": host_dev,
            "gpu_launches": int(launches),
            "stage_ms": {"vertex_clip_setup": float(stage_mean[0]), "binning": float(stage_mean[1]), "tile_raster_shade": tile_ms,
                         "geometry_to_resolve": float(stage_mean[3]), "frames_timed": 0 if stages is None else int(len(stages)),
                         "note": f"CUDA events around the stages of every {TIMING_STRIDE}th frame of the timed region"},
            "roofline": {"bound": "hbm", "kernel": "tile_kernel (tile raster + Forward+ shade + resolve + tonemap)",
                         "achieved": (b_tile / 1e9) / (tile_ms / 1e3) if tile_ms > 0 else None, "peak": hbm, "unit": "GB/s",
                         "frac": ((b_tile / 1e9) / (tile_ms / 1e3) / hbm) if tile_ms > 0 else None, "traffic": ncu_traffic()[0],
                         "traffic_source": ncu_traffic()[1], "algorithmic_bytes": b_tile, "peak_source": peak_src},
            "roofline_frame": {"algorithmic_bytes": b_frame, "achieved": (b_frame / 1e9) / (ms_dev / args.steps / 1e3),
                               "frac": (b_frame / 1e9) / (ms_dev / args.steps / 1e3) / hbm, "unit": "GB/s"},
            "clocks": clocks, "clocks_e2e": clocks_e2e,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_port(sd, args.cpu_seconds)
        emit(line)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


NSETS = 4


def workload_config(sd, world):
    """The `config` object of the JSON line -- the same dict from both arms (the driver compares them)."""
    tris = sum(len(sd.meshes[it["mesh"] - 1]["indices"]) // 3 for it in sd.items)
    n_lights = 0 if sd.lights is None else len(sd.lights)
    return {"workload": f"C2: 1080p Forward+ frame, {len(sd.items)} Suzanne instances ({tris} tris), {n_lights} point/spot lights, "
                        f"{int(sd.fp.tile_size)}-px tiles, <={int(sd.fp.max_lights_per_tile)} lights/tile"
                        + (f"; camera batch of {world}, one camera per GPU, LDR frames gathered to rank 0 over NCCL" if world > 1 else ""),
            "resolution": [W, H], "tile_size": int(sd.fp.tile_size), "max_lights_per_tile": int(sd.fp.max_lights_per_tile),
            "parallelism": f"sort-first camera batch x{world}" if world > 1 else "single GPU",
            "l2": f"{NSETS} rotating render-target sets ({NSETS * W * H * 24 / 1e6:.0f} MB) > 126 MB L2"}


def cpu_baseline_port(sd, budget_s):
    """The CPU oracle (port) on a bounded sample of the same workload: whole Forward+ frames, 1 thread."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import harness
    from oracle.bindings import Oracle
    o = Oracle("port")
    n, t0 = 0, time.perf_counter()
    while True:
        harness.cpu_forward(o, sd, forward_plus=True, aov=False)
        n += 1
        dt = time.perf_counter() - t0
        if dt >= budget_s or n >= 8:
            break
    return {"value": n / dt, "unit": "frames/s", "cores": 1, "kind": "port",
            "sample": f"{n} full C2 frames (light cull + Forward+ raster/shade + tonemap) by oracle/liboracle.so, single thread, {dt:.1f} s",
            "host_cores_available": os.cpu_count()}


def run_reference(args):
    """The reference's OWN CPU code for this path (oracle/_ref = its headers compiled from /root/reference):
    PassPBRForward::execute + PassTonemap::execute with ThreadPoolJobSystem(all host cores), as
    exp-plumbing/hello_pass_basics.cpp drives them.  NOTE the reference's CPU Forward+ lit pass shades the sun
    only (passes/pass_pbr_forward.hpp:157-195 never reads tile lists), so this arm does strictly LESS work per frame than the
    CUDA arm.  Its tile-list builder (cull_lights_tiled, lighting/jolt_light_culling.hpp:135-187 -- a serial loop in the reference)
    is part of the frame when oracle/_ref/libshs_lightcull_ref.so exists (the reference's header compiled against the JoltPhysics
    declaration shim, oracle/ref_lightcull_harness.cpp); the `sample` text says whether it was."""
    rank, _, world = dist_env()
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import harness
    from leisure_software_renderer_b200 import scenes
    from oracle import bindings
    kind = "reference" if bindings.available("reference") or os.path.isdir("/root/reference") else "port"
    o = bindings.Oracle(kind)
    cores = os.cpu_count() or 1
    o.set_threads(cores)
    sd = scenes.scene_c2(W, H)
    cull = None
    if kind == "reference":
        try:  # the reference's own tile-list builder over the frame's 1024 lights (A11), serial like in the reference
            import numpy as np
            from leisure_software_renderer_b200 import capi
            lref = bindings.LightCullReference()
            light_aabbs = np.ascontiguousarray(np.concatenate([sd.lights["cull_aabb_min"][:, :3], sd.lights["cull_aabb_max"][:, :3]], axis=1), dtype=np.float32)
            desc = capi.LightCullDesc(sd.viewproj, W, H, capi.LIGHT_CULL_TILED, 16, 128, z_near=sd.zn, z_far=sd.zf)
            lref.light_cull(light_aabbs, desc)
            cull = lambda: lref.light_cull(light_aabbs, desc)
        except Exception as e:  # library not built and no reference tree: the lit pass alone, as before
            print(f"reference arm: cull_lights_tiled unavailable ({e}); timing the lit pass only", file=sys.stderr)

    def frame():
        if kind == "reference":
            if cull is not None:
                cull()
            harness.cpu_forward(o, sd, forward_plus=False, aov=False)
        else:
            harness.cpu_forward(o, sd, forward_plus=True, aov=False)

    t0 = time.perf_counter(); frame(); t1 = time.perf_counter() - t0
    n_warm = max(args.warmup, 3)                                      # the CUDA arm's rule, so that both lines carry the same `warmup`
    steps = max(1, min(args.steps, int(150.0 / max(t1, 1e-3))))
    for _ in range(n_warm if t1 * n_warm < 30.0 else 1):
        frame()
    t0 = time.perf_counter()
    for _ in range(steps):
        frame()
    dt = time.perf_counter() - t0
    fps = steps / dt
    sample = (f"{steps} full C2 frames: " + ("cull_lights_tiled over 1024 lights (the reference's serial builder, Jolt declaration shim) + " if cull is not None else "") +
              f"PassPBRForward (sun + fake IBL only; the reference CPU path never shades tile light lists) + PassTonemap "
              f"via oracle/_ref/libshs_ref.so, ThreadPoolJobSystem({cores})" if kind == "reference"
              else f"{steps} full C2 Forward+ frames via oracle/liboracle.so (single thread)")
    line = {"impl": "reference", "metric": "frames/s", "value": fps, "unit": "frames/s", "n_gpus": world, "steps": steps,
            "steps_requested": args.steps, "warmup": n_warm, "ms_per_step": dt / steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(sd, world),
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores if kind == "reference" else 1, "kind": kind, "sample": sample},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    emit(line)


_JSON_OUT = None


def emit(line):
    print(json.dumps(line), file=_JSON_OUT or sys.stdout, flush=True)


def main():
    # stdout carries exactly ONE JSON line: everything else a library may print there (NCCL, torchrun children) goes to stderr
    global _JSON_OUT
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU-baseline sample budget")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
