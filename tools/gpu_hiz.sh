#!/bin/bash
# Hi-Z A/B: parity with the flag on, then C2 bench and C3/C4/C5 configs with SHSB_HIZ=0 / 1 on the same box
TAG=${1:-hiz}; OUT=gpurun_out/$TAG; mkdir -p $OUT
SHSB_HIZ=1 timeout 900 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_gather.py tests/test_gpu_sortfirst.py -x -q > $OUT/pytest_hiz.log 2>&1; echo "pytest(hiz=1) rc=$?"; tail -3 $OUT/pytest_hiz.log
for rep in 1 2; do for h in 0 1; do
  SHSB_HIZ=$h timeout 300 python bench.py --steps 300 --warmup 10 --no-cpu-baseline > $OUT/bench_hiz${h}_$rep.json 2> $OUT/bench_hiz${h}_$rep.err
  python - <<PY
import json
d=json.load(open("$OUT/bench_hiz${h}_$rep.json"))
print("C2 hiz=$h rep $rep: value", round(d["value"]), "tile alone ms", round(d["stage_ms"]["tile_raster_shade_alone"],4))
PY
done; done
for h in 0 1; do
  SHSB_HIZ=$h timeout 600 python tools/bench_configs.py c3 c4 c5 > $OUT/configs_hiz$h.jsonl 2> $OUT/configs_hiz$h.err
  python - <<PY
import json
for l in open("$OUT/configs_hiz$h.jsonl"):
    d=json.loads(l); print("hiz=$h", d["config"], "frame_ms", round(d["frame_ms"],3), "tile_ms", round(d.get("tile_ms",0),3))
PY
done
