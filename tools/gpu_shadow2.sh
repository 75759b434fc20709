#!/bin/bash
TAG=${1:-shadow2}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
SHSB_SHADOW_DIRECT=0 timeout 300 python tools/bench_shadow.py 30 > $OUT/bench_shadow_binned.json 2> $OUT/bench_shadow_binned.err; cat $OUT/bench_shadow_binned.json
timeout 300 python tools/bench_shadow.py 30 > $OUT/bench_shadow_direct.json 2> $OUT/bench_shadow_direct.err; cat $OUT/bench_shadow_direct.json
timeout 300 python tools/bench_shadow.py 30 2048 > $OUT/bench_shadow_direct_2048.json 2>> $OUT/bench_shadow_direct.err; cat $OUT/bench_shadow_direct_2048.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $OUT/launches.csv python tools/bench_shadow.py 6 > $OUT/ncu_launches.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:shadow_geometry_kernel -s 6 -c 1 -f -o $OUT/shadow_geometry_kernel python tools/bench_shadow.py 3 > $OUT/ncu_full.log 2>&1; echo "ncu full rc=$?"
python - <<PY
import csv, collections
rows=[r for r in csv.reader(open("$OUT/launches.csv")) if len(r)>5 and r[0].isdigit()]
agg=collections.defaultdict(list)
for r in rows: agg[r[4].split("(")[0][-40:]].append(float(r[-1]))
for k,v in agg.items(): print(f"{k:42s} n={len(v):3d} median {sorted(v)[len(v)//2]/1000:9.1f} us")
PY
timeout 300 python tools/bench_configs.py c3 > $OUT/config_c3.jsonl 2> $OUT/config_c3.err; cat $OUT/config_c3.jsonl | cut -c1-600
