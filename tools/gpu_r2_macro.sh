#!/bin/bash
TAG=${1:-r2macro}; OUT=gpurun_out/$TAG; mkdir -p $OUT
for v in macro4 macro6 macro12; do
  SHSB_LIB=$PWD/leisure_software_renderer_b200/libshsb_$v.so timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_light_bins.py tests/test_gpu_fuzz.py -m gpu -q -x -k "light or list or fplus or forward_plus or full_size_c2 or c5" > $OUT/pytest_$v.log 2>&1; echo "$v pytest rc=$?"; tail -1 $OUT/pytest_$v.log
done
bash tools/gpu_ab3.sh $TAG "default macro4 macro6 macro12" 2
