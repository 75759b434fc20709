#!/bin/bash
# Round-2 session A: baseline of the round-1 kernels under the fixed bench warm-up + first timings of L2 / L3 / scene culling.
TAG=${1:-r2a}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $OUT/smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 5 > $OUT/bench_20.json 2> $OUT/bench_20.err; echo "bench20 rc=$?"
timeout 300 python bench.py --steps 300 --warmup 10 --no-cpu-baseline > $OUT/bench_300.json 2> $OUT/bench_300.err; echo "bench300 rc=$?"
timeout 300 python tools/bench_configs.py c3 c4 c5 > $OUT/configs.jsonl 2> $OUT/configs.err; echo "configs rc=$?"
timeout 300 python tools/bench_legacy2.py 100 > $OUT/bench_legacy2.jsonl 2> $OUT/bench_legacy2.err; echo "legacy2 rc=$?"
timeout 300 python tools/bench_scene_cull.py 20 > $OUT/bench_scene_cull.jsonl 2> $OUT/bench_scene_cull.err; echo "scene cull rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:legacy2_ -s 12 -c 4 -f -o $OUT/legacy2_kernels \
    python tools/bench_legacy2.py 4 > $OUT/ncu_legacy2.log 2>&1; echo "ncu legacy2 rc=$?"
tail -3 $OUT/pytest_gpu.log; cat $OUT/bench_20.json; cat $OUT/bench_300.json
