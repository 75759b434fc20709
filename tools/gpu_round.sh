#!/bin/bash
# One GPU session: tests -> smoke -> bench -> reference arm -> ncu launch list -> ncu full capture of the tile kernel.
# Usage (on the GPU box, from the repo root): bash tools/gpu_round.sh <tag>
TAG=${1:-run}
OUT=gpurun_out/$TAG
mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/smoke.log
python bench.py --steps 300 --warmup 10 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 5 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/ncu_launches.log 2>&1; echo "ncu list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:tile_kernel -s 12 -c 1 -f -o $OUT/tile_kernel \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline > $OUT/ncu_full.log 2>&1; echo "ncu full rc=$?"
tail -3 $OUT/pytest_gpu.log; tail -2 $OUT/smoke.log; cat $OUT/bench.json
# legacy tile-job variant (configs[0] as shipped): frames/s next to the reference's threaded CPU path, and its two kernels under ncu
python tools/bench_legacy.py 300 > $OUT/bench_legacy.jsonl 2> $OUT/bench_legacy.err; echo "legacy bench rc=$?"
ncu --set full --clock-control none --import-source on -k regex:legacy_ -s 8 -c 2 -f -o $OUT/legacy_kernels \
    python tools/bench_legacy.py 20 > $OUT/ncu_legacy.log 2>&1; echo "ncu legacy rc=$?"
# legacy render-target demos (rows L2 / L3): parity + frames/s at the shipped sizes, and their kernels under ncu
python tools/bench_legacy2.py 100 > $OUT/bench_legacy2.jsonl 2> $OUT/bench_legacy2.err; echo "legacy2 bench rc=$?"
ncu --set full --clock-control none --import-source on -k regex:legacy2_ -s 12 -c 4 -f -o $OUT/legacy2_kernels \
    python tools/bench_legacy2.py 4 > $OUT/ncu_legacy2.log 2>&1; echo "ncu legacy2 rc=$?"
# scene-level culling calls (8f row 1): parity against the reference's own functions + calls/s through the C-ABI with host buffers
python tools/bench_scene_cull.py 20 > $OUT/bench_scene_cull.jsonl 2> $OUT/bench_scene_cull.err; echo "scene cull bench rc=$?"
# flat-shaded multi-light draws (8f row 1, the consumer of the light selections): parity + ms per frame next to the reference's own loop
python tools/bench_flat_draw.py 10 > $OUT/bench_flat_draw.jsonl 2> $OUT/bench_flat_draw.err; echo "flat draw bench rc=$?"
