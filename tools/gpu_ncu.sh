#!/bin/bash
# ncu --set full of the pass's kernels on one workload: gpu_ncu.sh TAG WORKLOAD [kernel regex]
tag=${1:-ncu}; wl=${2:-readme189}; kr=${3:-k_parse|k_lines|k_size|k_records|k_tile_scan|k_headers}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"$kr" -c ${4:-6} -s ${5:-12} -o gpurun_out/prof_${tag} python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-verify --workload $wl > gpurun_out/${tag}_ncu.log 2>&1
tail -2 gpurun_out/${tag}_ncu.log
