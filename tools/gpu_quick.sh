#!/bin/bash
# Quick GPU check: parity tests + one bench line.  Usage: bash tools/gpu_quick.sh <tag> [extra bench args]
TAG=${1:-q}; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_gpu.log
python bench.py --steps 300 --warmup 10 --no-cpu-baseline "$@" > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d = json.load(open("$OUT/bench.json"))
    print({k: d[k] for k in ("value", "ms_per_step", "stage_ms")}, d["e2e"]["value"], d["roofline"]["frac"])
except Exception as e:
    print("no bench line:", e); print(open("$OUT/bench.err").read()[-2000:])
PY
