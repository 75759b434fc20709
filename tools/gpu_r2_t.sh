#!/bin/bash
TAG=${1:-r2t}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 python -m pytest tests/test_zz_gpu_flat_draw.py -x -q -s > $OUT/pytest_gpu_flat.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/pytest_gpu_flat.log
