#!/bin/bash
# A/B of two builds of the library in ONE gpurun call (same box): bash tools/gpu_ab_lib.sh <tag> <old.so> [reps]
TAG=${1:-ab}; OLD=$2; REPS=${3:-2}; OUT=gpurun_out/$TAG; mkdir -p $OUT
for rep in $(seq 1 $REPS); do
  for which in new old; do
    if [ $which = old ]; then export SHSB_LIB=$PWD/$OLD; else unset SHSB_LIB; fi
    python bench.py --steps 400 --warmup 10 --no-cpu-baseline > $OUT/${which}_$rep.json 2> $OUT/${which}_$rep.err
    python - <<PY
import json
d=json.load(open("$OUT/${which}_$rep.json"))
print("$which rep $rep: value", round(d["value"]), "ms", round(d["ms_per_step"],4), "tile", round(d["stage_ms"]["tile_raster_shade"],4), "e2e", round(d["e2e"]["value"]))
PY
  done
done
