#!/bin/bash
tag=${1:-r2e}; n=${2:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/${tag}_bench_n$n.json 2> gpurun_out/${tag}_bench_n$n.err; echo "torchrun rc=$?"
tail -c 1500 gpurun_out/${tag}_bench_n$n.err
python bench.py --gpus $n --one-context --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_onectx_n$n.json 2> gpurun_out/${tag}_onectx_n$n.err; echo "onectx rc=$?"
tail -c 800 gpurun_out/${tag}_onectx_n$n.err
python - <<PY
import json
for f in ('gpurun_out/${tag}_bench_n$n.json','gpurun_out/${tag}_onectx_n$n.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value',round(d['value'],1),'e2e',d['e2e']['value'],d['e2e']['ms_per_step'],'stream',d['e2e_stream']['value'],'raw',d['raw_copy']['ms_per_step'],d['raw_copy']['GBps_per_gpu'],'verified',d['verified'])
        print('  onectx', d.get('one_context_all_gpus'))
    except Exception as e: print(f,'ERR',e)
PY
