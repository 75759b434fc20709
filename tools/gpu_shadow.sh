#!/bin/bash
TAG=${1:-shadow}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 300 python tools/bench_shadow.py 30 > $OUT/bench_shadow.json 2> $OUT/bench_shadow.err; echo "bench rc=$?"; cat $OUT/bench_shadow.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $OUT/launches.csv python tools/bench_shadow.py 6 > $OUT/ncu_launches.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tile_kernel -s 6 -c 1 -f -o $OUT/shadow_tile_kernel python tools/bench_shadow.py 3 > $OUT/ncu_full.log 2>&1; echo "ncu full rc=$?"
python - <<PY
import csv
rows=[r for r in csv.reader(open("$OUT/launches.csv")) if len(r)>5 and r[0].isdigit()]
import collections
agg=collections.defaultdict(list)
for r in rows: agg[r[4].split("(")[0][-40:]].append(float(r[-1]))
for k,v in agg.items(): print(f"{k:42s} n={len(v):3d} median {sorted(v)[len(v)//2]:9.1f} us")
PY
