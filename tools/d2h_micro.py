"""Which on-chip resource does a concurrent PCIe copy contend for?  Times three synthetic kernels (compute-bound matmul,
HBM-streaming copy, L2-resident gather) with and without a back-to-back 8.3 MB D2H / H2D copy loop on another stream."""
import time

import torch

dev = "cuda"
a = torch.randn(2048, 2048, device=dev, dtype=torch.bfloat16)
b = torch.randn(2048, 2048, device=dev, dtype=torch.bfloat16)
big_src = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
big_dst = torch.empty_like(big_src)
table = torch.randn(2 << 20, device=dev)                       # 8 MB, L2-resident
idx = torch.randint(0, table.numel(), (32 << 20,), device=dev)
fp32 = torch.randn(8 << 20, device=dev)
kernels = {
    "matmul bf16 2048^3 (tensor)": lambda: torch.matmul(a, b),
    "copy 256 MB (HBM)": lambda: big_dst.copy_(big_src),
    "gather 32M from 8 MB table (L2)": lambda: torch.index_select(table, 0, idx),
    "sin chain 8M fp32 (SIMT + MUFU)": lambda: torch.sin(torch.sin(torch.sin(fp32))),
}
copy = torch.cuda.Stream()
d_small = torch.empty(1920 * 1080 * 4, dtype=torch.uint8, device=dev)
h_small = [torch.empty(1920 * 1080 * 4, dtype=torch.uint8).pin_memory() for _ in range(2)]
main = torch.cuda.current_stream()
for name, fn in kernels.items():
    for mode in ("none", "d2h", "h2d"):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 40
        if mode != "none":
            with torch.cuda.stream(copy):
                for i in range(400):
                    if mode == "d2h":
                        h_small[i % 2].copy_(d_small, non_blocking=True)
                    else:
                        d_small.copy_(h_small[i % 2], non_blocking=True)
            time.sleep(0.002)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        e1.synchronize()
        busy = not copy.query()
        torch.cuda.synchronize()
        print(f"{name:36s} copy={mode:4s}: {e0.elapsed_time(e1) / n * 1e3:9.1f} us  (copy loop still running at the end: {busy})")
