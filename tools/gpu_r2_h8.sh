#!/bin/bash
# 16 x 8 raster tiles (libshsb_h8.so: every unit compiled with -DSHSB_TILE_H=8) against the default 16 x 16: parity suite on the variant, then an A/B of the bench and the configs
TAG=${1:-r2h8}; OUT=gpurun_out/$TAG; mkdir -p $OUT
SHSB_LIB=$PWD/leisure_software_renderer_b200/libshsb_h8.so timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu_h8.log 2>&1; echo "pytest(h8) rc=$?"; tail -5 $OUT/pytest_gpu_h8.log
bash tools/gpu_ab_lib.sh $TAG leisure_software_renderer_b200/libshsb_h8.so 2
for which in new old; do
  if [ $which = old ]; then export SHSB_LIB=$PWD/leisure_software_renderer_b200/libshsb_h8.so; else unset SHSB_LIB; fi
  timeout 600 python tools/bench_configs.py > $OUT/configs_$which.jsonl 2> $OUT/configs_$which.err
  python - <<PY
import json
for l in open("$OUT/configs_$which.jsonl"):
    d=json.loads(l); print("$which", d["config"], "frame_ms", round(d["frame_ms_min"],3), "tile_ms", round(d.get("tile_ms",0),3), "front", round(d.get("vertex_clip_setup_ms",0)+d.get("binning_ms",0),3))
PY
done
