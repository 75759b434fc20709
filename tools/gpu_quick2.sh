#!/bin/bash
# quick session: GPU suite + bench (300 steps) [+ optional extra command]
TAG=${1:-q}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
tail -5 $OUT/pytest_gpu.log
for n in ${STREAMS:-1 2}; do
  SHSB_TILE_STREAMS=$n timeout 300 python bench.py --steps 300 --warmup 10 --no-cpu-baseline > $OUT/bench_ts$n.json 2> $OUT/bench_ts$n.err; echo "bench ts=$n rc=$?"
  python - <<PY
import json
d=json.load(open("$OUT/bench_ts$n.json"))
print("streams $n: value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "tile_ms", d["stage_ms"]["tile_raster_shade_alone"], "host", d["host_submit_ms_per_step"]["total_ms"])
PY
done
