"""Per-kernel launch count / mean duration / share of a `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections
import csv
import sys


def main():
    lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
    rows = list(csv.DictReader(lines))
    agg = collections.OrderedDict()
    for r in rows:
        k = r["Kernel Name"].split("(")[0].split("::")[-1]
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += float(r["Metric Value"])
    tot = sum(v[1] for v in agg.values())
    print(f"| kernel | launches | mean us | share of GPU time |\n|---|---|---|---|")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {n} | {t / n / 1e3:.2f} | {t / tot * 100:.1f} % |")


if __name__ == "__main__":
    main()
