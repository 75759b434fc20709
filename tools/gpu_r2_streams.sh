#!/bin/bash
TAG=${1:-r2streams}; OUT=gpurun_out/$TAG; mkdir -p $OUT
for rep in 1 2; do for v in s2 s3 s4 cand2; do
  unset SHSB_LIB SHSB_TILE_STREAMS
  case $v in s2) export SHSB_TILE_STREAMS=2;; s3) export SHSB_TILE_STREAMS=3;; s4) export SHSB_TILE_STREAMS=4;; cand2) export SHSB_LIB=$PWD/leisure_software_renderer_b200/libshsb_cand2.so;; esac
  python bench.py --steps 400 --warmup 10 --no-cpu-baseline > $OUT/${v}_$rep.json 2> $OUT/${v}_$rep.err
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/${v}_$rep.json").read().strip().splitlines()[-1])
    print("$v rep $rep: value", round(d["value"]), "ms", round(d["ms_per_step"],4), "tile alone", round(d["stage_ms"]["tile_raster_shade_alone"],4), "e2e", round(d["e2e"]["value"]))
except Exception as e: print("$v ERR", e)
PY
done; done
