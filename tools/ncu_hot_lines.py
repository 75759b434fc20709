"""Top source lines of a kernel by executed instructions / stall samples: python tools/ncu_hot_lines.py rep [n]"""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    cur_file = "?"
    agg = {}
    hdr = None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            ie, ss = hdr.index("Instructions Executed"), hdr.index("# Samples")
            continue
        if hdr is None or r[0] in ("Function Name", "Kernel Name"):
            continue
        if r[2] != "-":  # SASS row
            continue
        try:
            key = (cur_file, int(r[0]), r[1].strip()[:110])
            inst, samp = int(r[ie] or 0), int(r[ss] or 0)
        except ValueError:
            continue
        a = agg.setdefault(key, [0, 0])
        a[0] += inst
        a[1] += samp
    tot_i = sum(v[0] for v in agg.values()) or 1
    tot_s = sum(v[1] for v in agg.values()) or 1
    print(f"total inst {tot_i}  samples {tot_s}  (first launch in the report only if several)")
    for (f, ln, src), (i, s) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:n]:
        print(f"{s / tot_s * 100:5.1f}% smp {i / tot_i * 100:5.1f}% inst  {f}:{ln:<4d} {src}")


if __name__ == "__main__":
    main()
