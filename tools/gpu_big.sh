#!/bin/bash
# N-GPU measurement of the named configs: default bench line, streamed configs 3 and 5, file->file CLI (N=1)
tag=${1:-r2f}; n=${2:-1}
mkdir -p gpurun_out
run() {  # run bench.py with the given args under torchrun when n > 1
  if [ $n -gt 1 ]; then python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n "$@"
  else python bench.py "$@"; fi
}
run --steps 20 --warmup 3 > gpurun_out/${tag}_readme_n$n.json 2> gpurun_out/${tag}_readme_n$n.err; echo "readme rc=$?"
run --workload short_reads --steps 10 --warmup 3 --stream-gb 20.4 --no-cpu-baseline > gpurun_out/${tag}_short_n$n.json 2> gpurun_out/${tag}_short_n$n.err; echo "short rc=$?"
run --workload contigs --per-gpu 40 --steps 10 --warmup 3 --stream-gb 25.4 --no-cpu-baseline > gpurun_out/${tag}_contigs_n$n.json 2> gpurun_out/${tag}_contigs_n$n.err; echo "contigs rc=$?"
for f in readme short contigs; do tail -c 300 gpurun_out/${tag}_${f}_n$n.err; done
if [ $n -eq 1 ]; then python tools/cli_time.py > gpurun_out/${tag}_cli.json 2> gpurun_out/${tag}_cli.err; cat gpurun_out/${tag}_cli.json; tail -3 gpurun_out/${tag}_cli.err; fi
python - <<PY
import json
for f in ('readme','short','contigs'):
    try:
        d=json.loads(open('gpurun_out/${tag}_%s_n$n.json'%f).read().strip().splitlines()[-1])
        print(f,'value',round(d['value'],1),'e2e',round(d['e2e']['value'],2),'stream',round(d['e2e_stream']['value'],2),'raw ms',round(d['raw_copy']['ms_per_step'],2),'e2e ms',round(d['e2e']['ms_per_step'],2),'verified',d['verified'])
        print('   one_ctx', d.get('one_context_all_gpus')); print('   stream', d.get('stream'))
    except Exception as e: print(f,'ERR',e)
PY
