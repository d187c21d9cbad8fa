#!/bin/bash
# A/B runs of library variants: gpu_ab.sh TAG "WORKLOADS" SPEC...   SPEC = name[:lib.so[:ENV=val,ENV=val]]
# prints one line per (spec, workload): device-resident Gbp/s, ms per step, per-kernel-group ms
tag=$1; wls=$2; shift 2
mkdir -p gpurun_out
for spec in "$@"; do
  IFS=: read -r name lib envs <<< "$spec"
  (
    if [ -n "$lib" ] && [ "$lib" != default ]; then export GOTRANSEQ_B200_LIB=$PWD/$lib; fi
    if [ -n "$envs" ]; then for kv in ${envs//,/ }; do export "$kv"; done; fi
    for wl in $wls; do
      extra="--no-verify"; if [ "$name" = verify ]; then extra=""; fi
      timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e $extra --workload $wl 2>>gpurun_out/${tag}_bench.err | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$name','$wl',round(d['value'],1),round(d['ms_per_step'],4),{k[:14]:round(x,4) for k,x in d['kernels_ms_per_step'].items()}, d.get('verified'))
except Exception as e: print('$name','$wl','ERR',e)"
    done
  )
done
tail -c 600 gpurun_out/${tag}_bench.err
