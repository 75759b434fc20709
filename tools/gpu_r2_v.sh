#!/bin/bash
TAG=${1:-r2v}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests/test_zz_gpu_legacy2.py tests/test_zz_gpu_golden_round1b.py tests/test_gpu_legacy.py -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_gpu.log
timeout 300 python tools/bench_legacy2.py > $OUT/bench_legacy2.jsonl 2> $OUT/bench_legacy2.err; cut -c1-330 $OUT/bench_legacy2.jsonl; tail -3 $OUT/bench_legacy2.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_legacy2.csv python tools/bench_legacy2.py 3 > $OUT/ncu.log 2>&1; echo "ncu rc=$?"
python - <<PY
import csv,collections
rows=[r for r in csv.reader(open("$OUT/launches_legacy2.csv")) if len(r)>10]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value")
agg=collections.OrderedDict()
for r in rows[1:]:
    agg.setdefault(r[ki][:60],[]).append(float(r[vi].replace(",","")))
for k,v in agg.items(): print(k, len(v), "launches, mean us", sum(v)/len(v)/1000.0 if max(v)>1000 else sum(v)/len(v), "max", max(v))
PY
