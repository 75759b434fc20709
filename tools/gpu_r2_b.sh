#!/bin/bash
# Round-2 session B: render streams -- GPU suite + bench at 1..4 render streams.
TAG=${1:-r2b}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
tail -5 $OUT/pytest_gpu.log
for n in 1 2 3 4; do
  SHSB_TILE_STREAMS=$n timeout 300 python bench.py --steps 300 --warmup 10 --no-cpu-baseline > $OUT/bench_ts$n.json 2> $OUT/bench_ts$n.err; echo "bench ts=$n rc=$?"
  python - <<PY
import json
d=json.load(open("$OUT/bench_ts$n.json"))
print("streams $n: value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "tile_ms", d["stage_ms"]["tile_raster_shade"], "host", d["host_submit_ms_per_step"]["total_ms"])
PY
done
SHSB_TILE_STREAMS=2 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench_ts2_s20.json 2> $OUT/bench_ts2_s20.err; cat $OUT/bench_ts2_s20.json | python -c "import json,sys; d=json.load(sys.stdin); print('steps20: value', round(d['value']), 'e2e', round(d['e2e']['value']))"
