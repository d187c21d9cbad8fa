#!/bin/bash
# e2e A/B: gpu_e2e.sh TAG SPEC...   SPEC = name[:lib.so|default[:ENV=val,ENV=val]]
tag=$1; shift
mkdir -p gpurun_out
for spec in "$@"; do
  IFS=: read -r name lib envs <<< "$spec"
  (
    if [ -n "$lib" ] && [ "$lib" != default ]; then export GOTRANSEQ_B200_LIB=$PWD/$lib; fi
    if [ -n "$envs" ]; then for kv in ${envs//,/ }; do export "$kv"; done; fi
    python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-verify 2>gpurun_out/${tag}_${name}.err | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$name', round(d['value'],1), 'e2e', round(d['e2e']['value'],2), round(d['e2e']['ms_per_step'],3), 'stream', round(d['e2e_stream']['value'],2), round(d['e2e_stream']['ms_per_step'],3), 'stream1', d['e2e_stream'].get('single_thread_reader'), 'raw', round(d['raw_copy']['ms_per_step'],3))"
  )
done
