"""Writes a markdown summary of one kernel's `ncu --set full` capture:
python tools/ncu_report_md.py capture.ncu-rep title > profiles/xxx.md"""
import csv
import io
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def run(*a):
    return subprocess.run([sys.executable, *a], capture_output=True, text=True).stdout


def main():
    rep, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else sys.argv[1])
    print(f"# {title}\n")
    print(f"Source: `{os.path.basename(rep)}` (`ncu --set full --clock-control none --import-source on`, one launch, captured under gpurun;")
    print("times under ncu are cold-cache and serialised: use them for shares and counters, not as bench values).\n")
    print("## Key counters\n```")
    print(run(os.path.join(HERE, "ncu_summary.py"), rep).strip())
    print("```\n## Warp stall samples (pc sampling)\n```")
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, d = rows[0], rows[2]
    st = [(h[i].replace("smsp__pcsamp_warps_issue_stalled_", ""), int(float(d[i]))) for i in range(len(h))
          if h[i].startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in h[i]]
    tot = sum(v for _, v in st) or 1
    for k, v in sorted(st, key=lambda kv: -kv[1]):
        if v:
            print(f"{k:24s} {v:6d}  {v / tot * 100:5.1f}%")
    print("```\n## Executed warp-instructions and stall samples by code region\n```")
    print(run(os.path.join(HERE, "ncu_regions.py"), rep).strip())
    print("```\n## Hottest source lines (by stall samples)\n```")
    print(run(os.path.join(HERE, "ncu_hot_lines.py"), rep, "30").strip())
    print("```")


if __name__ == "__main__":
    main()
