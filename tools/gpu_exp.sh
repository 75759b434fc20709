#!/bin/bash
# bench variants selected by environment variables: bash tools/gpu_exp.sh <tag> "VAR=val ..." "VAR=val ..." ...
TAG=$1; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
i=0
for envs in "$@"; do
  i=$((i+1))
  env $envs python bench.py --steps 300 --warmup 10 --no-cpu-baseline > $OUT/bench_$i.json 2> $OUT/bench_$i.err
  python - <<PY
import json
try:
    d = json.load(open("$OUT/bench_$i.json"))
    print("[$envs]", round(d["value"]), "fps", round(d["ms_per_step"]*1e3,1), "us |", {k: round(v*1e3,1) for k, v in d["stage_ms"].items() if k != "frames_timed"}, "| e2e", round(d["e2e"]["value"]))
except Exception as e:
    print("[$envs] failed", e); print(open("$OUT/bench_$i.err").read()[-1500:])
PY
done
