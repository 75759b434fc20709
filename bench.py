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
    """Samples SM clocks / throttle reasons during the timed region (B200_PROFILING.md recipe).  The driver's 20-step run times a
    3 ms region, far below nvidia-smi's sampling period, so the samples come from NVML directly (nvidia_ml_py: one query is ~0.1 ms)
    on a helper thread, about one per millisecond; nvidia-smi is the fall-back.  `post_roll` keeps sampling while the caller
    runs more (untimed) steps under the same load when the region itself was too short for a handful of samples."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None
        self.samples = []          # (sm_mhz, reasons bitmask) from NVML
        self.nvml = None
        self._stop = threading.Event()
        self.in_region = 0
        try:
            import pynvml
            pynvml.nvmlInit()
            # CUDA_VISIBLE_DEVICES-relative ordinal -> NVML handle through the CUDA device's UUID when torch can tell it
            try:
                import torch
                uuid = str(torch.cuda.get_device_properties(gpu_index).uuid)
                self.handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode() if not uuid.startswith("GPU-") else uuid.encode())
            except Exception:
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _nvml_loop(self):
        n = self.nvml
        while not self._stop.is_set():
            try:
                mhz = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                try:
                    reasons = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                except Exception:
                    reasons = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                self.samples.append((mhz, reasons))
            except Exception:
                break
            time.sleep(0.0005)

    def start(self):
        if self.nvml is not None:
            self.t = threading.Thread(target=self._nvml_loop, daemon=True)
            self.t.start()
            return
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

    def n_samples(self) -> int:
        return len(self.samples) if self.nvml is not None else len(self.lines)

    def mark_region_start(self):
        """Samples taken before this point (the warm-up) are dropped."""
        if self.nvml is not None:
            del self.samples[:]
        else:
            self.lines.clear()

    def mark_region_end(self):
        self.in_region = self.n_samples()

    def stop(self) -> dict:
        if self.nvml is not None:
            self._stop.set()
            self.t.join(timeout=1.0)
            n = self.nvml
            names = (("hw_slowdown", getattr(n, "nvmlClocksEventReasonHwSlowdown", 0x8)), ("hw_thermal_slowdown", getattr(n, "nvmlClocksEventReasonHwThermalSlowdown", 0x40)),
                     ("sw_thermal_slowdown", getattr(n, "nvmlClocksEventReasonSwThermalSlowdown", 0x20)), ("sw_power_cap", getattr(n, "nvmlClocksEventReasonSwPowerCap", 0x4)))
            reasons = sorted({name for _, r in self.samples for name, bit in names if r & bit})
            sm = [m for m, _ in self.samples]
            return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": len(sm),
                    "samples_in_timed_region": self.in_region, "source": "NVML, ~1 sample / ms; samples beyond the timed region were taken while the same steps kept running"}
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
                "reasons": sorted(reasons), "samples": len(sm), "samples_in_timed_region": self.in_region, "source": "nvidia-smi -lms 20"}


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
    batch = world > 1
    if batch:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # NCCL chatter (e.g. its version line) must not share stdout with the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    ctx = Context(local_rank)
    base = scenes.scene_c2(W, H)
    # N = 1: the C2 frame, one camera, a step = one frame.  N > 1: BASELINE configs[4]'s 64-camera batch over the same scene, a step = the
    # whole batch (fixed total work => strong scaling), rank r renders cameras {c : c mod N == r} and pushes every LDR frame into the
    # assembly memory on rank 0's GPU (shsb_frame_gather: copy-engine peer writes over NVLink, no collective on the data path).
    if batch:
        ring = scenes.camera_ring(base, N_CAMERAS)
        from leisure_software_renderer_b200 import sortfirst
        my_cams = sortfirst.cameras_of_rank(N_CAMERAS, world, rank)
        views = [ring[c] for c in my_cams]
    else:
        my_cams, views = [0], [base]
    frames_per_step = N_CAMERAS if batch else 1
    sd = views[0]
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
    frame_bytes = W * H * 4
    host_ldr = [torch.empty(frame_bytes, dtype=torch.uint8).pin_memory() for _ in range(4)]

    gather = None
    gstream = None
    if batch:
        # the root allocates 2 slots x 64 frames of assembly memory; its CUDA IPC export travels as bytes (an NCCL broadcast of a uint8 tensor)
        exp_t = torch.zeros(C_sizeof_export(capi), dtype=torch.uint8, device="cuda")
        if rank == 0:
            gather, exp = ctx.gather_create(world, GATHER_SLOTS, N_CAMERAS * frame_bytes)
            exp_t.copy_(torch.frombuffer(bytearray(bytes(exp)), dtype=torch.uint8))
        dist.broadcast(exp_t, src=0)
        if rank != 0:
            gather = ctx.gather_open(bytes(exp_t.cpu().numpy().tobytes()), rank)
        gstream = torch.cuda.ExternalStream(ctx.gather_stream(), device=local_rank)
    counters = {"step": 0, "frame": 0}

    def step(_i, e2e=False):
        counters["step"] += 1
        s_no = counters["step"]
        if e2e:
            ctx.lights_upload(lights_pinned.numpy())          # H2D 160 B x n_lights from pinned memory, once per step
        for c, v in zip(my_cams, views):
            k = counters["frame"] % NSETS
            counters["frame"] += 1
            hdr, dm, ldr = sets[k]
            ctx.frame_forward_plus(v.scene, fp, hdr, dm, ldr, want_stats=False)   # asynchronous; draw list H2D inside the call
            if batch:
                ctx.frame_gather(gather, s_no, ldr, capi.PLANE_COLOR, 0, frame_bytes, c * frame_bytes)   # frame assembly over NVLink, behind the frame
            if e2e:
                # D2H of the frame into pinned memory (copy stream: overlaps with the next frames' rendering)
                ctx.rt_download_async(ldr, capi.PLANE_COLOR, host_ldr[counters["frame"] % 4].data_ptr(), frame_bytes)
        if batch:
            ctx.gather_commit(gather, s_no)
            if rank == 0:
                ctx.gather_wait(gather, s_no)                 # the assembled batch is complete on the root's GPU ...
                ctx.gather_release(gather, s_no)              # ... and its slot may be reused two steps later

    def barrier():
        torch.cuda.synchronize()
        ctx.sync()
        if batch:
            dist.barrier()
        torch.cuda.synchronize()

    # one frame with statistics (also sizes the per-frame arena), then warm-up
    st = ctx.frame_forward_plus(sd.scene, fp, *sets[0]).as_dict()
    counts, _ = ctx.light_lists_download()
    n_warm = max(args.warmup, 3)

    def warm(e2e):
        """Warm-up of the path that is timed next: the same calls as the timed loop (both executable-graph topologies of the
        front end -- with and without stage events --, the light-upload ring, the read-back streams and their pinned buffers, the
        assembly ring), enough frames that every render-target set and every transient arena has been through it once."""
        if not e2e:
            ctx.timing_enable(TIMING_STRIDE)
        for i in range(n_warm if batch else max(n_warm, 2 * NSETS)):
            step(i, e2e)
        if e2e:
            ctx.sync()
        barrier()
        if not e2e:
            ctx.timing_collect()
            ctx.timing_enable(False)

    def timed(e2e):
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()          # the helper thread is up and sampling before the region starts (the warm-up runs under the same load)
        warm(e2e)
        launches0 = ctx.launch_count()
        barrier()
        e0, e1, e1g = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        if not e2e:
            ctx.timing_enable(TIMING_STRIDE)   # stage events on every 8th frame only: they are extra commands on the critical stream
        ctx.host_submit_us(reset=True)
        if rank == 0:
            sampler.mark_region_start()
        e0.record(stream)
        t_host = time.perf_counter()
        for i in range(args.steps):
            step(i, e2e)
        host_ms = (time.perf_counter() - t_host) * 1e3 / args.steps   # host time to SUBMIT a step (includes back-pressure waits on the staging ring)
        hu = ctx.host_submit_us(reset=True)
        host_ms = {"total_ms": host_ms, "library_us": {k: float(hu[i] / max(hu[5], 1.0)) for i, k in
                   enumerate(("draw_list", "staging_incl_ring_wait", "arena_checks", "capture_enqueue", "graph_update_launch_tile_launch"))}}
        if e2e:
            ctx.sync()  # the last read-backs (copy stream) belong to the timed region
        ctx.fence()     # frames run on several render streams: the main stream (and e1) behind all of them
        e1.record(stream)
        if batch:
            e1g.record(gstream)  # the last pushes / the root's wait for every rank's last commit belong to the timed region
        barrier()
        ms = e0.elapsed_time(e1)
        if batch:
            ms = max(ms, e0.elapsed_time(e1g))
        stages = ctx.timing_collect() if not e2e else None
        if not e2e:
            ctx.timing_enable(False)
        if rank == 0:
            sampler.mark_region_end()
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if batch:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        launches = ctx.launch_count() - launches0
        # A 20-step region lasts a few milliseconds: keep the same steps running (untimed) for about 60 ms so that the sampler sees
        # the load for more than a handful of samples.  The count is derived from the all-reduced time: every rank runs the same
        # number of steps (a batch step ends in the root's wait for every rank's commit).
        n_roll = int(min(2000, max(8, 60.0 / max(float(t.item()) / args.steps, 1e-3))))
        for i in range(n_roll):
            step(i, e2e)
        barrier()
        clocks = sampler.stop() if rank == 0 else None
        return float(t.item()), stages, clocks, launches, host_ms

    ms_dev, stages, clocks, launches, host_dev = timed(False)
    ms_e2e, _, clocks_e2e, _, host_e2e = timed(True)

    # The tile kernel timed ALONE for the roofline: every tile kernel on the main stream, back to back (front ends of later frames
    # still overlap it, as in round 1), CUDA events around each launch.  In the timed regions above tile kernels of consecutive
    # frames overlap on two render streams, so an event pair around one of them also spans its neighbour's share of the SMs.
    ctx.set_tile_streams(1)
    ctx.timing_enable(2)
    for i in range(48):
        ctx.frame_forward_plus(sd.scene, fp, *sets[i % NSETS], want_stats=False)
    alone = ctx.timing_collect()
    ctx.timing_enable(False)
    ctx.set_tile_streams(2)
    tile_alone_ms = float(np.median(alone[4:, 2])) if len(alone) > 8 else float("nan")

    n_items = len(sd.items)
    n_blocks = sum((len(sd.meshes[it["mesh"] - 1]["indices"]) // 3 + 127) // 128 for it in sd.items)
    h2d = frames_per_step * (n_items * 256 + n_blocks * 8) + world * len(sd.lights) * 160
    d2h = frames_per_step * frame_bytes
    fps = frames_per_step * args.steps / (ms_dev / 1e3)
    fps_e2e = frames_per_step * args.steps / (ms_e2e / 1e3)
    hbm, peak_src = peaks()
    mesh = sd.meshes[0]
    b_frame, b_tile = algorithmic_bytes(sd, counts, len(mesh["positions"]), len(mesh["indices"]))
    stage_mean = stages.mean(axis=0) if stages is not None and len(stages) else np.zeros(4)
    tile_ms = float(stage_mean[2])

    if rank == 0:
        line = {
            "metric": "frames/s", "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": n_warm,
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "strong" if batch else "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": workload_config(base, world),
            "frames_per_step": frames_per_step,
            "mtri_per_s": fps * st["tri_input"] / 1e6, "mfrag_per_s": fps * st["frag_covered"] / 1e6,
            "frame_stats": st,
            "e2e": {"value": fps_e2e, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps, "host_submit_ms_per_step": host_e2e},
            "host_submit_ms_per_step": host_dev,
            "gpu_launches": int(launches),
            "stage_ms": {"vertex_clip_setup": float(stage_mean[0]), "binning": float(stage_mean[1]), "tile_raster_shade_overlapped": tile_ms,
                         "tile_raster_shade_alone": tile_alone_ms,
                         "geometry_to_resolve": float(stage_mean[3]), "frames_timed": 0 if stages is None else int(len(stages)),
                         "note": f"CUDA events around the stages of every {TIMING_STRIDE}th frame of the timed region (tile kernels of consecutive frames "
                                 f"overlap on 2 render streams there); `alone` = median over {max(0, len(alone) - 4)} launches of a separate pass with every "
                                 f"tile kernel on one stream"},
            "roofline": {"bound": "hbm", "kernel": "tile_kernel (tile raster + Forward+ shade + resolve + tonemap)",
                         "achieved": (b_tile / 1e9) / (tile_alone_ms / 1e3) if tile_alone_ms > 0 else None, "peak": hbm, "unit": "GB/s",
                         "frac": ((b_tile / 1e9) / (tile_alone_ms / 1e3) / hbm) if tile_alone_ms > 0 else None, "traffic": ncu_traffic()[0],
                         "traffic_source": ncu_traffic()[1], "algorithmic_bytes": b_tile, "peak_source": peak_src,
                         "launch_ms": tile_alone_ms, "how": "kernel timed alone on its stream (CUDA events around each launch, back-to-back frames)"},
            "roofline_frame": {"algorithmic_bytes": b_frame, "achieved": (b_frame / 1e9) / (ms_dev / args.steps / frames_per_step * world / 1e3),
                               "frac": (b_frame / 1e9) / (ms_dev / args.steps / frames_per_step * world / 1e3) / hbm, "unit": "GB/s",
                               "note": "whole frame (front end + tile kernel, pipelined) per GPU: algorithmic bytes / steady-state period"},
            "clocks": clocks, "clocks_e2e": clocks_e2e,
        }
        if not batch and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_port(sd, args.cpu_seconds)
        emit(line)
    if gather is not None:
        barrier()
        ctx.gather_destroy(gather)
    ctx.close()
    if batch:
        dist.destroy_process_group()


def C_sizeof_export(capi):
    import ctypes
    return ctypes.sizeof(capi.GatherExport)


NSETS = 4
N_CAMERAS = 64     # BASELINE configs[4]: the 64-camera batch
GATHER_SLOTS = 2


def workload_config(sd, world):
    """The `config` object of the JSON line -- the same dict from both arms (the driver compares them)."""
    tris = sum(len(sd.meshes[it["mesh"] - 1]["indices"]) // 3 for it in sd.items)
    n_lights = 0 if sd.lights is None else len(sd.lights)
    frame = (f"1080p Forward+ frame, {len(sd.items)} Suzanne instances ({tris} tris), {n_lights} point/spot lights, "
             f"{int(sd.fp.tile_size)}-px tiles, <={int(sd.fp.max_lights_per_tile)} lights/tile")
    if world > 1:
        workload = (f"C5b: {N_CAMERAS}-camera batch (cameras on a ring around the C2 scene), each camera one {frame}; a step = the whole batch; sort-first: "
                    f"cameras c mod {world} per GPU, every LDR frame assembled on rank 0's GPU over NVLink (copy-engine peer pushes, shsb_frame_gather)")
    else:
        workload = "C2: " + frame
    return {"workload": workload, "resolution": [W, H], "tile_size": int(sd.fp.tile_size), "max_lights_per_tile": int(sd.fp.max_lights_per_tile),
            "parallelism": f"sort-first camera batch, {N_CAMERAS} cameras over {world} GPUs" if world > 1 else "single GPU",
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
    base = scenes.scene_c2(W, H)
    # N = 1: the C2 frame.  N > 1: the CUDA arm's workload is the 64-camera batch; a reference step is a bounded sample of it -- one
    # camera of the ring per step, in ring order (64 full frames per step would take 16 s of CPU each).
    views = scenes.camera_ring(base, N_CAMERAS) if world > 1 else [base]
    sd = views[0]
    cull = None
    if kind == "reference":
        try:  # the reference's own tile-list builder over the frame's 1024 lights (A11), serial like in the reference
            import numpy as np
            from leisure_software_renderer_b200 import capi
            lref = bindings.LightCullReference()
            light_aabbs = np.ascontiguousarray(np.concatenate([sd.lights["cull_aabb_min"][:, :3], sd.lights["cull_aabb_max"][:, :3]], axis=1), dtype=np.float32)
            descs = [capi.LightCullDesc(v.viewproj, W, H, capi.LIGHT_CULL_TILED, 16, 128, z_near=v.zn, z_far=v.zf) for v in views]
            lref.light_cull(light_aabbs, descs[0])
            cull = lambda k: lref.light_cull(light_aabbs, descs[k])
        except Exception as e:  # library not built and no reference tree: the lit pass alone, as before
            print(f"reference arm: cull_lights_tiled unavailable ({e}); timing the lit pass only", file=sys.stderr)
    frame_no = [0]

    def frame():
        k = frame_no[0] % len(views)
        frame_no[0] += 1
        if kind == "reference":
            if cull is not None:
                cull(k)
            harness.cpu_forward(o, views[k], forward_plus=False, aov=False)
        else:
            harness.cpu_forward(o, views[k], forward_plus=True, aov=False)

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
    sample = ((f"{steps} frames = one camera of the {N_CAMERAS}-camera batch per step, in ring order: " if world > 1 else f"{steps} full C2 frames: ") + ("cull_lights_tiled over 1024 lights (the reference's serial builder, Jolt declaration shim) + " if cull is not None else "") +
              f"PassPBRForward (sun + fake IBL only; the reference CPU path never shades tile light lists) + PassTonemap "
              f"via oracle/_ref/libshs_ref.so, ThreadPoolJobSystem({cores})" if kind == "reference"
              else f"{steps} full C2 Forward+ frames via oracle/liboracle.so (single thread)")
    line = {"impl": "reference", "metric": "frames/s", "value": fps, "unit": "frames/s", "n_gpus": world, "steps": steps,
            "steps_requested": args.steps, "warmup": n_warm, "ms_per_step": dt / steps * 1e3, "higher_is_better": True,
            "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(base, world),
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
