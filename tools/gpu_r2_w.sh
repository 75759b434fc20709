#!/bin/bash
# ncu --set full of the legacy render-target demos' lit raster kernel (the Suzanne draw of the soft-shadow demo: launch 2 of the kernel) and the direct shadow kernel
TAG=${1:-r2w}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:legacy2_raster_kernel -s 4 -c 1 -o $OUT/legacy2_raster -f python tools/bench_legacy2.py 3 > $OUT/ncu_raster.log 2>&1; echo "ncu raster rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:legacy2_shadow_direct_kernel -s 4 -c 1 -o $OUT/legacy2_shadow -f python tools/bench_legacy2.py 3 > $OUT/ncu_shadow.log 2>&1; echo "ncu shadow rc=$?"
ls -la $OUT
