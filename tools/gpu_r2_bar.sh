#!/bin/bash
TAG=${1:-r2bar}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_pipeline.py tests/test_gpu_fuzz.py tests/test_gpu_sortfirst.py -m gpu -q -x 2>&1 | tail -1
bash tools/gpu_ab3.sh $TAG "default old" 2
for rep in 1 2; do for v in default old; do
  if [ $v = default ]; then unset SHSB_LIB; else export SHSB_LIB=$PWD/leisure_software_renderer_b200/libshsb_$v.so; fi
  timeout 300 python tools/bench_configs.py c3 c4 c5 > $OUT/configs_${v}_$rep.jsonl 2> $OUT/configs.err; python -c "
import json
for l in open('$OUT/configs_${v}_$rep.jsonl'):
    d=json.loads(l); print('$v', d['config'], 'frame', round(d['frame_ms_min'],3), 'tile', round(d['tile_ms'],3))"
done; done
