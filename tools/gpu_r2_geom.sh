#!/bin/bash
TAG=${1:-r2geom}; OUT=gpurun_out/$TAG; mkdir -p $OUT
for rep in 1 2; do for v in default geom8 geom10 geom12; do
  if [ $v = default ]; then unset SHSB_LIB; else export SHSB_LIB=$PWD/leisure_software_renderer_b200/libshsb_$v.so; fi
  timeout 300 python tools/bench_configs.py c4 c5 > $OUT/configs_${v}_$rep.jsonl 2> $OUT/configs.err; python -c "
import json
for l in open('$OUT/configs_${v}_$rep.jsonl'):
    d=json.loads(l); print('$v', d['config'], 'frame', round(d['frame_ms_min'],3), 'geom', round(d['vertex_clip_setup_ms'],3), 'bin', round(d['binning_ms'],3), 'tile', round(d['tile_ms'],3))"
done; done
bash tools/gpu_ab3.sh $TAG "default geom8 geom10" 1
