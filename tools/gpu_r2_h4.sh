#!/bin/bash
TAG=${1:-r2h4}; OUT=gpurun_out/$TAG; mkdir -p $OUT
SHSB_LIB=$PWD/leisure_software_renderer_b200/libshsb_h4.so timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_pipeline.py tests/test_gpu_fuzz.py -m gpu -q -x > $OUT/pytest_gpu_h4.log 2>&1; echo "pytest(h4) rc=$?"; tail -3 $OUT/pytest_gpu_h4.log
for rep in 1 2; do for v in h4 h8 h16; do
  if [ $v = h16 ]; then unset SHSB_LIB; else export SHSB_LIB=$PWD/leisure_software_renderer_b200/libshsb_$v.so; fi
  python bench.py --steps 400 --warmup 10 --no-cpu-baseline > $OUT/${v}_$rep.json 2> $OUT/${v}_$rep.err
  python - <<PY
import json
d=json.loads(open("$OUT/${v}_$rep.json").read().strip().splitlines()[-1])
print("$v rep $rep: value", round(d["value"]), "ms", round(d["ms_per_step"],4), "tile alone", round(d["stage_ms"]["tile_raster_shade_alone"],4), "e2e", round(d["e2e"]["value"]))
PY
done; done
for v in h4 h8 h16; do
  if [ $v = h16 ]; then unset SHSB_LIB; else export SHSB_LIB=$PWD/leisure_software_renderer_b200/libshsb_$v.so; fi
  timeout 600 python tools/bench_configs.py > $OUT/configs_$v.jsonl 2> $OUT/configs_$v.err
  python - <<PY
import json
for l in open("$OUT/configs_$v.jsonl"):
    d=json.loads(l); print("$v", d["config"], "frame_ms", round(d["frame_ms_min"],3), "tile_ms", round(d.get("tile_ms",0),3), "front", round(d.get("vertex_clip_setup_ms",0)+d.get("binning_ms",0),3))
PY
done
