#!/bin/bash
tag=${1:-r2c}
lib=${2:-}
[ -n "$lib" ] && export GOTRANSEQ_B200_LIB=$PWD/$lib
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'k_parse|k_lines|k_size|k_records|k_tile_scan|k_headers' -c 6 -s 12 -o gpurun_out/prof_${tag} python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${tag}_ncu.log 2>&1
tail -2 gpurun_out/${tag}_ncu.log
