#!/bin/bash
# 8K sort-first strong scaling on one multi-GPU box.  Usage: bash tools/gpu_sortfirst.sh <tag>
TAG=${1:-sf}; OUT=gpurun_out/$TAG; mkdir -p $OUT
NG=$(nvidia-smi -L | wc -l)
timeout 300 python tools/bench_sortfirst.py > $OUT/n1.json 2> $OUT/n1.err; echo "N=1 rc=$?"; tail -c 600 $OUT/n1.json
for N in 2 4 8; do
  [ $N -gt $NG ] && break
  for L in bands interleaved; do
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700 + N)) tools/bench_sortfirst.py --layout $L > $OUT/n${N}_$L.json 2> $OUT/n${N}_$L.err
    echo "N=$N $L rc=$?"; grep "^{" $OUT/n${N}_$L.json | python -c "import sys,json; [print({k:d[k] for k in ('value','ms_per_frame','tile_rows_per_rank','frag_shaded_per_rank','tri_input_per_rank','assembled_frame_equals_whole_frame')}) for d in map(json.loads, sys.stdin)]" || tail -5 $OUT/n${N}_$L.err
  done
done
