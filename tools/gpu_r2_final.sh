#!/bin/bash
# Round-2 closing profile session (16 x 8 raster tiles): whole GPU suite, smoke, bench (driver settings and long), reference arm, launch list, full capture of the tile
# kernel, the other configs, the side benches
TAG=${1:-r2final}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?"
timeout 300 python bench.py --steps 20 --warmup 5 > $OUT/bench_20.json 2> $OUT/bench_20.err; echo "bench20 rc=$?"
timeout 300 python bench.py --steps 300 --warmup 10 --no-cpu-baseline > $OUT/bench_300.json 2> $OUT/bench_300.err; echo "bench300 rc=$?"
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/ncu_launches.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tile_kernel -s 12 -c 1 -f -o $OUT/tile_kernel python bench.py --steps 6 --warmup 3 --no-cpu-baseline > $OUT/ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 300 python tools/bench_configs.py c3 c4 c5 > $OUT/configs.jsonl 2> $OUT/configs.err; echo "configs rc=$?"
timeout 300 python tools/bench_legacy2.py 100 > $OUT/bench_legacy2.jsonl 2> $OUT/bench_legacy2.err; echo "legacy2 rc=$?"
timeout 300 python tools/bench_legacy.py 300 > $OUT/bench_legacy.jsonl 2> $OUT/bench_legacy.err; echo "legacy rc=$?"
timeout 300 python tools/bench_scene_cull.py 20 > $OUT/bench_scene_cull.jsonl 2> $OUT/bench_scene_cull.err; echo "scene cull rc=$?"
timeout 300 python tools/bench_flat_draw.py 10 > $OUT/bench_flat_draw.jsonl 2> $OUT/bench_flat_draw.err; echo "flat draw rc=$?"
python - <<PY
import json
for f in ("bench_20","bench_300"):
    d=json.loads(open("$OUT/"+f+".json").read().strip().splitlines()[-1]); print(f, "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "roofline", d["roofline"], "clocks", d["clocks"])
PY
