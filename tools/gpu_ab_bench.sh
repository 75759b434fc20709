mkdir -p gpurun_out/s4i
for cfg in "8 0" "1 0" "1000000 0"; do set -- $cfg;
  SHSB_BENCH_STRIDE=$1 python bench.py --steps 300 --warmup 10 --no-cpu-baseline > gpurun_out/s4i/b_$1_$2.json 2> gpurun_out/s4i/b_$1_$2.err
  python - <<PY
import json
d=json.load(open("gpurun_out/s4i/b_$1_$2.json"))
print("stride $1 noclocks $2: value", round(d["value"]), "ms", round(d["ms_per_step"],4), "tile", round(d["stage_ms"]["tile_raster_shade"],4), "e2e", round(d["e2e"]["value"]), d["clocks"])
PY
done
