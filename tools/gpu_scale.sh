#!/bin/bash
# Scaling run on one multi-GPU box: bench.py at N = 1, 2, 4, 8 (as many as the box has).  Usage: bash tools/gpu_scale.sh <tag> [steps]
TAG=${1:-scale}; STEPS=${2:-200}
OUT=gpurun_out/$TAG
mkdir -p $OUT
NG=$(nvidia-smi -L | wc -l)
for N in 1 2 4 8; do
  [ $N -gt $NG ] && break
  if [ $N -eq 1 ]; then
    timeout 300 python bench.py --gpus 1 --steps $STEPS --warmup 5 --no-cpu-baseline > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) \
      bench.py --gpus $N --steps $STEPS --warmup 5 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err
  fi
  echo "N=$N rc=$?"
  python - <<PY
import json
try:
    d = [json.loads(l) for l in open("$OUT/bench_n$N.json") if l.startswith("{")][-1]
    print("  value", round(d["value"]), "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), "e2e ms/step", round(d["e2e"]["ms_per_step"], 4))
except Exception as e:
    print("  no line:", e); print(open("$OUT/bench_n$N.err").read()[-1500:])
PY
done
