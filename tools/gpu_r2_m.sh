#!/bin/bash
TAG=${1:-r2m}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 python -m pytest tests/test_zz_gpu_scene_cull.py -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
timeout 300 python tools/bench_occlusion.py 10 > $OUT/bench_occlusion.jsonl 2> $OUT/bench_occlusion.err; cat $OUT/bench_occlusion.jsonl
