#!/bin/bash
TAG=${1:-r2fast4}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_pipeline.py tests/test_gpu_fuzz.py tests/test_gpu_post_passes.py -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $OUT/pytest_gpu.log
for v in fast general; do
  if [ $v = general ]; then export SHSB_NO_FAST_TILE=1; else unset SHSB_NO_FAST_TILE; fi
  for rep in 1 2; do
  timeout 300 python tools/bench_configs.py c3 c4 c5 > $OUT/configs_${v}_$rep.jsonl 2> $OUT/configs.err; python -c "
import json
for l in open('$OUT/configs_${v}_$rep.jsonl'):
    d=json.loads(l); print('$v', d['config'], 'frame', round(d['frame_ms_min'],3), 'tile', round(d['tile_ms'],3))"
  done
done
