"""Pinned-memory D2H / H2D bandwidth of this box for the e2e leg's copy sizes (8.3 MB LDR frame)."""
import subprocess
import torch

print(subprocess.run(["nvidia-smi", "--query-gpu=pcie.link.gen.current,pcie.link.gen.max,pcie.link.width.current", "--format=csv"], capture_output=True, text=True).stdout)
for nbytes in (8294400, 33177600, 268435456):
    d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    for name, src, dst in (("D2H", d, h), ("H2D", h, d)):
        for _ in range(3):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 50
        e0.record()
        for _ in range(n):
            dst.copy_(src, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        print(f"{name} {nbytes / 1e6:8.1f} MB: {ms * 1e3:8.1f} us  {nbytes / ms / 1e6:6.1f} GB/s")
