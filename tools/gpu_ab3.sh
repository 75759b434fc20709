#!/bin/bash
# A/B of library builds on one box: bash tools/gpu_ab3.sh <tag> "<variant> <variant> ..." [reps]   (variant "default" = libshsb.so, else libshsb_<variant>.so)
TAG=${1:-ab3}; VARS=${2:-"default"}; REPS=${3:-2}; OUT=gpurun_out/$TAG; mkdir -p $OUT
for rep in $(seq 1 $REPS); do for v in $VARS; do
  if [ $v = default ]; then unset SHSB_LIB; else export SHSB_LIB=$PWD/leisure_software_renderer_b200/libshsb_$v.so; fi
  python bench.py --steps 400 --warmup 10 --no-cpu-baseline > $OUT/${v}_$rep.json 2> $OUT/${v}_$rep.err
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/${v}_$rep.json").read().strip().splitlines()[-1])
    print("$v rep $rep: value", round(d["value"]), "ms", round(d["ms_per_step"],4), "tile alone", round(d["stage_ms"]["tile_raster_shade_alone"],4), "e2e", round(d["e2e"]["value"]))
except Exception as e: print("$v rep $rep ERR", e)
PY
done; done
