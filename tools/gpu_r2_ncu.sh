#!/bin/bash
# Round-2 profile session: bench (driver settings and long), launch list, full capture of the tile kernel, L2 / L3 / scene-cull timings
TAG=${1:-r2ncu}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 300 python bench.py --steps 20 --warmup 5 > $OUT/bench_20.json 2> $OUT/bench_20.err; echo "bench20 rc=$?"
timeout 300 python bench.py --steps 300 --warmup 10 --no-cpu-baseline > $OUT/bench_300.json 2> $OUT/bench_300.err; echo "bench300 rc=$?"
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/ncu_launches.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tile_kernel -s 12 -c 1 -f -o $OUT/tile_kernel python bench.py --steps 6 --warmup 3 --no-cpu-baseline > $OUT/ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 300 python tools/bench_legacy2.py 100 > $OUT/bench_legacy2.jsonl 2> $OUT/bench_legacy2.err; echo "legacy2 rc=$?"
timeout 300 python tools/bench_scene_cull.py 20 > $OUT/bench_scene_cull.jsonl 2> $OUT/bench_scene_cull.err; echo "scene cull rc=$?"
timeout 300 python tools/bench_configs.py c3 c4 c5 > $OUT/configs.jsonl 2> $OUT/configs.err; echo "configs rc=$?"
python - <<PY
import json
for f in ("bench_20","bench_300"):
    d=json.load(open("$OUT/"+f+".json")); print(f, "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "roofline", round(d["roofline"]["frac"],4), "clocks", d["clocks"])
PY
