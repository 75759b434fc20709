"""Prints a compact summary of an .ncu-rep (raw page): python tools/ncu_summary.py file.ncu-rep [regex]"""
import csv
import io
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "launch__waves_per_multiprocessor",
    "launch__grid_size", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.sum", "sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__sass_average_branch_targets_threads_uniform.pct",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__cycles_active.avg", "sm__cycles_elapsed.max",
    "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__average_warp_latency_issue_stalled",
]


def main():
    rep = sys.argv[1]
    extra = sys.argv[2] if len(sys.argv) > 2 else None
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    print("kernels:", [r[hdr.index("Kernel Name")][:50] for r in data])
    for i, h in enumerate(hdr):
        short = h.split(".", 2)[-1] if h.count(".") >= 2 and h.split(".")[1] in ("TriageCompute",) else h
        if any(short == k or h == k for k in KEYS) or "warp_issue_stalled" in h and "per_warp_active" in h or (extra and re.search(extra, h)):
            print(f"{h[:95]:95s} {units[i]:10s} " + "  ".join(r[i] for r in data))


if __name__ == "__main__":
    main()
