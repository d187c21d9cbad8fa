#!/bin/bash
# A/B of the end-to-end legs (one-shot, streaming with 4 / 1 reader threads) between library builds
# usage: gpu_stream_ab.sh TAG SPEC...   SPEC = name[:lib.so[:extra bench.py flag]]
tag=$1; shift
mkdir -p gpurun_out
for spec in "$@"; do
  IFS=: read -r name lib extra <<< "$spec"
  ( if [ -n "$lib" ]; then export GOTRANSEQ_B200_LIB=$PWD/$lib; fi
    timeout 120 python bench.py --steps 5 --warmup 3 --e2e-steps 10 --no-cpu-baseline --no-verify $extra 2>>gpurun_out/${tag}.err | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$name','e2e',round(d['e2e']['ms_per_step'],2),'stream4',round(d['e2e_stream']['ms_per_step'],2),'stream1',round(d['e2e_stream']['single_thread_reader']['ms_per_step'],2),'raw',round(d['raw_copy']['ms_per_step'],2))" )
done
