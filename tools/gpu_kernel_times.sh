#!/bin/bash
# per-kernel solo durations (ncu launch list) of the C2 bench loop for one library build: bash tools/gpu_kernel_times.sh <tag> [SHSB_LIB path]
TAG=$1; LIB=$2
OUT=gpurun_out/$TAG; mkdir -p $OUT
SHSB_LIB=$LIB ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/launches.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/ncu.log 2>&1
python tools/launch_shares.py $OUT/launches.csv
