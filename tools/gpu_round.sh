#!/bin/bash
# One GPU session: parity suites (each under its own timeout), then the bench line.  Logs in gpurun_out/.
tag=${1:-r2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader > gpurun_out/${tag}_gpu.txt 2>&1
nproc >> gpurun_out/${tag}_gpu.txt; free -g | head -2 >> gpurun_out/${tag}_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/${tag}_summary.txt
timeout 600 python -m pytest tests/test_engine.py -m gpu -q -x --timeout 240 > gpurun_out/${tag}_test_engine.log 2>&1; echo "engine rc=$?" | tee -a gpurun_out/${tag}_summary.txt
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_split.py -m gpu -q --timeout 300 > gpurun_out/${tag}_test_parity.log 2>&1; echo "parity rc=$?" | tee -a gpurun_out/${tag}_summary.txt
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?" | tee -a gpurun_out/${tag}_summary.txt
tail -3 gpurun_out/${tag}_test_engine.log gpurun_out/${tag}_test_parity.log
tail -c 600 gpurun_out/${tag}_bench.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${tag}_bench.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','verified','kernels_ms_per_step')})
    print('e2e',d['e2e']['value'],d['e2e']['ms_per_step'],'stream',d['e2e_stream']['value'],d['e2e_stream']['ms_per_step'],'raw',d['raw_copy']['ms_per_step'])
except Exception as e: print('no bench line',e)
PY
