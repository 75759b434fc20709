#!/bin/bash
# ncu launch list + full captures of the kernels added in session 4 (post passes, cell cull).  Usage: bash tools/gpu_ncu_new.sh <tag>
TAG=${1:-new}; OUT=gpurun_out/$TAG; mkdir -p $OUT
python tools/bench_post.py > $OUT/post.jsonl 2> $OUT/post.err; echo "post rc=$?"
python tools/bench_post.py bins > $OUT/bins.jsonl 2> $OUT/bins.err; echo "bins rc=$?"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum --clock-control none -c 300 --csv --log-file $OUT/post_launches.csv python tools/bench_post.py > $OUT/ncu_post.log 2>&1; echo "ncu post rc=$?"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:"cell_cull|frustum_list|tile_depth_range|macro_cull|tile_cull" -c 60 --csv --log-file $OUT/bins_launches.csv python tools/bench_post.py bins > $OUT/ncu_bins.log 2>&1; echo "ncu bins rc=$?"
cat $OUT/post.jsonl $OUT/bins.jsonl
