#!/bin/bash
tag=${1:-r2g}; n=${2:-1}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_split.py -m gpu -q -x --timeout 300 > gpurun_out/${tag}_split_tests.log 2>&1; echo "split tests rc=$?"; tail -3 gpurun_out/${tag}_split_tests.log
python bench.py --workload chromosome_split --steps 10 --warmup 3 --verify > gpurun_out/${tag}_split_n1.json 2> gpurun_out/${tag}_split_n1.err; echo "split bench rc=$?"; tail -c 400 gpurun_out/${tag}_split_n1.err
python -c "
import json; d=json.loads(open('gpurun_out/${tag}_split_n1.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['verified'], d['gpu_launches'])"
