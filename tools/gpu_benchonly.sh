#!/bin/bash
tag=${1:-b}; shift
mkdir -p gpurun_out
for v in default "$@"; do
  if [ $v = default ]; then unset GOTRANSEQ_B200_LIB; else export GOTRANSEQ_B200_LIB=$PWD/$v; fi
  for wl in readme189 contigs chromosome short_reads; do
    timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-verify --workload $wl 2>>gpurun_out/${tag}_bench.err | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v','$wl',round(d['value'],1),round(d['ms_per_step'],4),{k[:14]:round(x,4) for k,x in d['kernels_ms_per_step'].items()})"
  done
done
