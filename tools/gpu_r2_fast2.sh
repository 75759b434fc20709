#!/bin/bash
TAG=${1:-r2fast3}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -s > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -a "PSNR fast vs general\|passed\|failed" $OUT/pytest_gpu.log | tail -6
for rep in 1 2; do for v in fast general; do
  if [ $v = general ]; then export SHSB_NO_FAST_TILE=1; else unset SHSB_NO_FAST_TILE; fi
  python bench.py --steps 400 --warmup 10 --no-cpu-baseline > $OUT/${v}_$rep.json 2> $OUT/${v}_$rep.err
  python - <<PY
import json
d=json.loads(open("$OUT/${v}_$rep.json").read().strip().splitlines()[-1])
print("$v rep $rep: value", round(d["value"]), "ms", round(d["ms_per_step"],4), "tile alone", round(d["stage_ms"]["tile_raster_shade_alone"],4), "e2e", round(d["e2e"]["value"]))
PY
done; done
unset SHSB_NO_FAST_TILE
timeout 300 python tools/bench_configs.py c5 > $OUT/configs_c5.jsonl 2> $OUT/configs.err; python -c "
import json
for l in open('$OUT/configs_c5.jsonl'):
    d=json.loads(l); print(d['config'], 'frame', round(d['frame_ms_min'],3), 'tile', round(d['tile_ms'],3))"
