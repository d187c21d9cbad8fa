#!/bin/bash
# the round's last GPU session: the whole -m gpu suite, smoke(), the default bench line
tag=${1:-last}
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -q -x --timeout 300 > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/${tag}_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"
timeout 200 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/${tag}_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','verified','kernels_ms_per_step','clocks','gpu_launches')})
print('e2e',d['e2e']['value'],'stream',d['e2e_stream']['value'],'raw',d['raw_copy']['value'])
print('roofline',d['roofline']['frac'],'path',d['path']['frac_of_peak'], 'cpu', d.get('cpu_baseline',{}).get('value'))
PY
