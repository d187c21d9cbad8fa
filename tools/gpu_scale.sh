#!/bin/bash
# the named configs on N GPUs of one box: bash tools/gpu_scale.sh TAG N
tag=${1:-r2h}; n=${2:-2}
bash tools/gpu_big.sh $tag $n
if [ $n -gt 1 ]; then
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n --workload chromosome_split --steps 10 --warmup 3 --verify > gpurun_out/${tag}_split_n$n.json 2> gpurun_out/${tag}_split_n$n.err
else
  python bench.py --workload chromosome_split --steps 10 --warmup 3 --verify > gpurun_out/${tag}_split_n$n.json 2> gpurun_out/${tag}_split_n$n.err
fi
echo "split rc=$?"; tail -c 300 gpurun_out/${tag}_split_n$n.err
python -c "
import json; d=json.loads(open('gpurun_out/${tag}_split_n$n.json').read().strip().splitlines()[-1]); print('split n=$n', d['value'], d['ms_per_step'], d['verified'], d['segments_cover_output'])"
