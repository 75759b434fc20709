#!/bin/bash
TAG=${1:-r2u}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests/test_zz_gpu_legacy2.py tests/test_zz_gpu_golden_round1b.py tests/test_zz_gpu_scene_cull.py tests/test_zz_gpu_flat_draw.py -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 $OUT/pytest_gpu.log
timeout 300 python tools/bench_flat_draw.py 10 > $OUT/bench_flat_draw.jsonl 2> $OUT/bench_flat_draw.err; cat $OUT/bench_flat_draw.jsonl; tail -3 $OUT/bench_flat_draw.err
timeout 300 python tools/bench_scene_cull.py > $OUT/bench_scene_cull.jsonl 2> $OUT/bench_scene_cull.err; cut -c1-260 $OUT/bench_scene_cull.jsonl; tail -3 $OUT/bench_scene_cull.err
timeout 300 python tools/bench_legacy2.py > $OUT/bench_legacy2.jsonl 2> $OUT/bench_legacy2.err; cut -c1-300 $OUT/bench_legacy2.jsonl; tail -3 $OUT/bench_legacy2.err
