"""Aggregate pinned-memory D2H bandwidth of the box when N GPUs copy an 8.3 MB LDR frame to the host at the same time
(the e2e leg of bench.py at N > 1: every rank reads its own camera's frame back over its own PCIe link).
Usage: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/pcie_bw_multi.py"""
import json
import os

import torch
import torch.distributed as dist

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
nbytes = 1920 * 1080 * 4
d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
h = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in range(2)]
for i in range(5):
    h[i % 2].copy_(d, non_blocking=True)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
n = 300
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(n):
    h[i % 2].copy_(d, non_blocking=True)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
t = torch.tensor([ms], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    worst = float(t.item())
    print(json.dumps({"n_gpus": world, "bytes": nbytes, "us_per_copy_max_over_ranks": worst * 1e3, "per_gpu_gbs": nbytes / worst / 1e6,
                      "aggregate_gbs": world * nbytes / worst / 1e6}))
if world > 1:
    dist.destroy_process_group()
