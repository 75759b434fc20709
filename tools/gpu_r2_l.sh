#!/bin/bash
TAG=${1:-r2l}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
timeout 300 python tools/bench_occlusion.py 10 > $OUT/bench_occlusion.jsonl 2> $OUT/bench_occlusion.err; cat $OUT/bench_occlusion.jsonl
timeout 300 python tools/bench_shadow.py 30 > $OUT/bench_shadow.json 2> $OUT/bench_shadow.err; cat $OUT/bench_shadow.json
timeout 300 python tests/cpp/_build/drop_in_test > $OUT/drop_in_test.log 2>&1; tail -3 $OUT/drop_in_test.log
