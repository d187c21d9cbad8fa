#!/bin/bash
# pipeline probe (tools/pipeline_probe.py) on the shapes given; usage: gpu_pipe.sh TAG "WORKLOADS"
tag=${1:-pipe}; wls=${2:-readme189}
mkdir -p gpurun_out
for wl in $wls; do
  timeout 240 python tools/pipeline_probe.py $wl 2>>gpurun_out/${tag}.err | tee -a gpurun_out/${tag}.log
done
tail -c 400 gpurun_out/${tag}.err
