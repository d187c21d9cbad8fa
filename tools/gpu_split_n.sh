#!/bin/bash
tag=${1:-r2i}; n=${2:-2}
mkdir -p gpurun_out
if [ $n -gt 1 ]; then
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n --workload chromosome_split --steps 20 --warmup 5 --verify > gpurun_out/${tag}_split_n$n.json 2> gpurun_out/${tag}_split_n$n.err
else
  python bench.py --workload chromosome_split --steps 20 --warmup 5 --verify > gpurun_out/${tag}_split_n$n.json 2> gpurun_out/${tag}_split_n$n.err
fi
echo "split rc=$?"; grep -v "^\s*$" gpurun_out/${tag}_split_n$n.err | tail -5 | cut -c1-300
python -c "
import json; d=json.loads(open('gpurun_out/${tag}_split_n$n.json').read().strip().splitlines()[-1]); print('split n=$n', d['value'], d['ms_per_step'], d['verified'], d['segments_cover_output'], d['gpu_launches'])"
