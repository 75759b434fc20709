#!/bin/bash
# 8K sort-first strong scaling on one multi-GPU box.  Usage: bash tools/gpu_sortfirst.sh <tag>
TAG=${1:-sf}; OUT=gpurun_out/$TAG; mkdir -p $OUT
NG=$(nvidia-smi -L | wc -l)
show() { grep "^{" $1 | python -c "import sys,json; [print({k:d[k] for k in ('n_gpus','layout','value','ms_per_frame','tile_rows_per_rank','frag_shaded_per_rank','tri_input_per_rank','host_submit_ms_per_rank','library_host_ms_per_rank','gpu_stage_ms_per_rank','frame_assembly_ms_per_rank','assembled_frame_equals_whole_frame')}) for d in map(json.loads, sys.stdin)]" || tail -5 ${1%.json}.err; }
timeout 300 python tools/bench_sortfirst.py --steps 30 > $OUT/n1.json 2> $OUT/n1.err; echo "N=1 rc=$?"; show $OUT/n1.json
for cfg in "8 bands" "4 bands" "2 bands"; do
  set -- $cfg; N=$1; L=$2
  [ $N -gt $NG ] && continue
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700 + N)) tools/bench_sortfirst.py --layout $L --steps 30 > $OUT/n${N}_$L.json 2> $OUT/n${N}_$L.err
  echo "N=$N $L rc=$?"; show $OUT/n${N}_$L.json
done
