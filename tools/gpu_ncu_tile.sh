#!/bin/bash
# ncu --set full capture of one tile_kernel launch in the C2 bench loop.  Usage: bash tools/gpu_ncu_tile.sh <tag>
TAG=${1:-ncu}; OUT=gpurun_out/$TAG; mkdir -p $OUT
python bench.py --steps 6 --warmup 3 --no-cpu-baseline > $OUT/plain.log 2>&1 || { echo "bench failed"; tail -5 $OUT/plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:tile_kernel -s 12 -c 1 -f -o $OUT/tile_kernel \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline > $OUT/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la $OUT
