#!/bin/bash
tag=${1:-r2b}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_split.py tests/test_engine.py -m gpu -q -x --timeout 300 > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/${tag}_tests.log
for v in default p5 p8; do
  if [ $v = default ]; then unset GOTRANSEQ_B200_LIB; else export GOTRANSEQ_B200_LIB=$PWD/build/lib_$v.so; fi
  for wl in readme189 contigs short_reads; do
    timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --workload $wl 2>>gpurun_out/${tag}_bench.err | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v','$wl',round(d['value'],1),round(d['ms_per_step'],4),{k[:14]:round(x,4) for k,x in d['kernels_ms_per_step'].items()})"
  done
done
