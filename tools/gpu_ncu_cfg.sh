#!/bin/bash
# ncu --set full of one kernel inside tools/bench_configs.py: bash tools/gpu_ncu_cfg.sh <tag> <config> <kernel-regex> [skip]
TAG=$1; CFG=$2; KER=$3; SKIP=${4:-2}
OUT=gpurun_out/$TAG; mkdir -p $OUT
ncu --set full --clock-control none --import-source on -k regex:$KER -s $SKIP -c 1 -f -o $OUT/${KER}_$CFG \
    python tools/bench_configs.py $CFG --iters 2 > $OUT/ncu_${KER}_$CFG.log 2>&1; echo "ncu rc=$?"; tail -2 $OUT/ncu_${KER}_$CFG.log | cut -c1-300
