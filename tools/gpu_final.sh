#!/bin/bash
# End-of-round evidence on one GPU: bench line, reference arm, ncu launch list, ncu --set full of the
# big kernels, file->file CLI timing.  Everything lands in gpurun_out/ with the given tag.
tag=${1:-r2z}
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2>> gpurun_out/${tag}_bench.err; echo "ref rc=$?"
for wl in contigs chromosome short_reads; do
  python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu-baseline >> gpurun_out/${tag}_workloads.jsonl 2>> gpurun_out/${tag}_bench.err; echo "$wl rc=$?"
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-verify > gpurun_out/${tag}_ncu_list.log 2>&1; echo "ncu list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'k_parse|k_lines' -c 2 -s 10 -o gpurun_out/prof_${tag}_full python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-verify > gpurun_out/${tag}_ncu_full.log 2>&1; echo "ncu full rc=$?"
python tools/cli_time.py > gpurun_out/${tag}_cli.json 2> gpurun_out/${tag}_cli.err; echo "cli rc=$?"; cat gpurun_out/${tag}_cli.json
python - <<PY
import json
d=json.loads(open('gpurun_out/${tag}_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','verified','kernels_ms_per_step','clocks','gpu_launches')})
print('e2e',d['e2e']['value'],'stream',d['e2e_stream']['value'],d['e2e_stream'].get('single_thread_reader'),'raw',d['raw_copy']['value'])
print('roofline',d['roofline']['frac'],'path',d['path']['frac_of_peak'], 'cpu', d.get('cpu_baseline',{}).get('value'))
PY
