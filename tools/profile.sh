#!/bin/bash
# Run on the GPU box (under gpurun).  (1) launch list of a short bench.py run (every kernel with its device time),
# (2) one --set full capture of every GEMM / attention / LayerNorm / embedding launch of tools/step_once.py.
# Each ncu command runs only after the same command line exited 0 without ncu.   Usage: bash tools/profile.sh <tag>
set -u
tag=${1:-r01}
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_$tag.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_$tag.csv $CMD > gpurun_out/ncu_list_$tag.log 2>&1
CMD2="python tools/step_once.py"
$CMD2 > gpurun_out/plain2_$tag.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'gemm_bf16_tc|attn_fwd_bf16|add_ln|embed_compose' -c 400 -f -o gpurun_out/step_$tag $CMD2 > gpurun_out/ncu_full_$tag.log 2>&1
tail -n 2 gpurun_out/ncu_list_$tag.log | cut -c1-300; tail -n 3 gpurun_out/ncu_full_$tag.log
