#!/bin/bash
# Upper bound of what a different store mechanism (TMA) could gain: the tile kernel with its stores predicated off vs the normal build
TAG=${1:-storebound}; OUT=gpurun_out/$TAG; mkdir -p $OUT
for rep in 1 2; do for which in normal nostore; do
  if [ $which = nostore ]; then export SHSB_LIB=$PWD/leisure_software_renderer_b200/libshsb_nostore.so; else unset SHSB_LIB; fi
  timeout 300 python bench.py --steps 300 --warmup 10 --no-cpu-baseline > $OUT/bench_${which}_$rep.json 2> $OUT/bench_${which}_$rep.err
  python - <<PY
import json
d=json.load(open("$OUT/bench_${which}_$rep.json"))
print("$which rep $rep: value", round(d["value"]), "tile alone ms", round(d["stage_ms"]["tile_raster_shade_alone"],4), "tile overlapped ms", round(d["stage_ms"]["tile_raster_shade_overlapped"],4))
PY
done; done
