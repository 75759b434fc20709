#!/bin/bash
# bench.py at N = 1 and N = $1.. (one box with >= N GPUs): usage bash tools/gpu_scale2.sh <tag> "1 2 4 8" [steps]
TAG=${1:-scale}; NS=${2:-"1 2"}; STEPS=${3:-20}
OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt
for n in $NS; do
  if [ $n = 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps $STEPS --warmup 5 --no-cpu-baseline > $OUT/bench_n1.json 2> $OUT/bench_n1.err; echo "n=1 rc=$?"
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --steps $STEPS --warmup 5 > $OUT/bench_n$n.json 2> $OUT/bench_n$n.err; echo "n=$n rc=$?"
  fi
  python - <<PY
import json
try:
    d=json.loads([l for l in open("$OUT/bench_n$n.json") if l.strip().startswith("{")][-1])
    print("N=$n value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"],4), "roofline", d["roofline"]["frac"], "tile alone", d["stage_ms"]["tile_raster_shade_alone"])
except Exception as e:
    print("N=$n: no line", e); import subprocess; print(subprocess.run(["tail","-30","$OUT/bench_n$n.err"],capture_output=True,text=True).stdout)
PY
done
