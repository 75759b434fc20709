#!/bin/bash
# checkpoint: whole GPU suite, smoke, the bench as the driver runs it (20 steps) and with 300 steps, the reference arm
TAG=${1:-r2y}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_gpu.log
python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke.log
python bench.py --steps 20 --warmup 5 > $OUT/bench_20.json 2> $OUT/bench_20.err; echo "bench20 rc=$?"
python bench.py --steps 300 --warmup 10 > $OUT/bench_300.json 2> $OUT/bench_300.err; echo "bench300 rc=$?"
python bench.py --impl reference --steps 5 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref rc=$?"
python - <<PY
import json
for f in ("bench_20","bench_300","bench_ref"):
    try:
        d=json.loads(open("$OUT/"+f+".json").read().strip().splitlines()[-1])
        print(f, "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms/step", round(d["ms_per_step"],4), "roofline", d.get("roofline",{}).get("frac"), "clocks", d.get("clocks",{}).get("sm_mhz"), d.get("clocks",{}).get("reasons"))
    except Exception as e: print(f, "ERR", e)
PY
