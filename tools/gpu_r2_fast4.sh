#!/bin/bash
TAG=${1:-r2fast5}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $OUT/pytest_gpu.log
for rep in 1 2; do for v in fast general; do
  if [ $v = general ]; then export SHSB_NO_FAST_TILE=1; else unset SHSB_NO_FAST_TILE; fi
  python bench.py --steps 400 --warmup 10 --no-cpu-baseline > $OUT/${v}_$rep.json 2> $OUT/${v}_$rep.err
  python - <<PY
import json
d=json.loads(open("$OUT/${v}_$rep.json").read().strip().splitlines()[-1])
print("$v rep $rep: value", round(d["value"]), "ms", round(d["ms_per_step"],4), "tile alone", round(d["stage_ms"]["tile_raster_shade_alone"],4), "e2e", round(d["e2e"]["value"]))
PY
done; done
