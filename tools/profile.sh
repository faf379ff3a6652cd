#!/bin/bash
# Run on the GPU box (under gpurun): launch list of a short bench, then one --set full capture of the GEMM kernel.
# Usage: bash tools/profile.sh <tag>
set -u
tag=${1:-r01}
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_$tag.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_$tag.csv $CMD > gpurun_out/ncu_list_$tag.log 2>&1
$CMD > gpurun_out/plain2_$tag.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tc -s 120 -c 8 -f -o gpurun_out/gemm_$tag $CMD > gpurun_out/ncu_full_$tag.log 2>&1
tail -3 gpurun_out/ncu_list_$tag.log gpurun_out/ncu_full_$tag.log
