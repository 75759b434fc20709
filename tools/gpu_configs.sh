#!/bin/bash
# bash tools/gpu_configs.sh <tag> [configs...]
TAG=$1; shift
timeout 800 python tools/bench_configs.py --iters 12 "$@" > gpurun_out/configs_$TAG.jsonl 2> gpurun_out/configs_$TAG.err; echo rc=$?; tail -5 gpurun_out/configs_$TAG.err
python - <<PY
import json
for l in open("gpurun_out/configs_$TAG.jsonl"):
    d = json.loads(l); print(d["config"], round(d["frame_ms"],3), "ms", {k: (round(v,3) if isinstance(v,float) else v) for k,v in d.items() if k.endswith("_ms") and k!="frame_ms"}, d["stats"], d["properties"], "hbm_frac", round(d["hbm_frac"],3), "Mtri/s", round(d["mtri_per_s"]))
PY
