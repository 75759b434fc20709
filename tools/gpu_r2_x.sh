#!/bin/bash
TAG=${1:-r2x}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests/test_zz_gpu_legacy2.py tests/test_zz_gpu_golden_round1b.py tests/test_gpu_legacy.py -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_gpu.log
timeout 300 python tools/bench_legacy2.py > $OUT/bench_legacy2.jsonl 2> $OUT/bench_legacy2.err; cut -c1-330 $OUT/bench_legacy2.jsonl; tail -3 $OUT/bench_legacy2.err
timeout 300 python tools/bench_legacy.py > $OUT/bench_legacy.jsonl 2> $OUT/bench_legacy.err; cut -c1-400 $OUT/bench_legacy.jsonl; tail -3 $OUT/bench_legacy.err
