#!/bin/bash
TAG=${1:-r2nl}; OUT=gpurun_out/$TAG; mkdir -p $OUT
for rep in 1 2; do for v in default nl10 nl12; do
  if [ $v = default ]; then unset SHSB_LIB; else export SHSB_LIB=$PWD/leisure_software_renderer_b200/libshsb_$v.so; fi
  timeout 300 python tools/bench_configs.py c3 c4 > $OUT/configs_${v}_$rep.jsonl 2> $OUT/configs.err; python -c "
import json
for l in open('$OUT/configs_${v}_$rep.jsonl'):
    d=json.loads(l); print('$v', d['config'], 'frame', round(d['frame_ms_min'],3), 'tile', round(d['tile_ms'],3))"
done; done
