#!/bin/bash
# Run on the GPU box (under gpurun): launch list of a short bench, then --set full captures of the GEMM and the
# attention kernel.  Each ncu command runs only after the same command line exited 0 without ncu.
# Usage: bash tools/profile.sh <tag>
set -u
tag=${1:-r01}
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_$tag.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/launches_$tag.csv $CMD > gpurun_out/ncu_list_$tag.log 2>&1
# full capture of the kernels of ONE eager step (the roofline leg at the end of the bench launches them eagerly)
$CMD > gpurun_out/plain2_$tag.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'gemm_bf16_tc|attn_fwd_bf16|add_ln|embed_compose' --launch-skip-before-match 0 -s 7000 -c 90 -f -o gpurun_out/step_$tag $CMD > gpurun_out/ncu_full_$tag.log 2>&1
tail -n 3 gpurun_out/ncu_list_$tag.log gpurun_out/ncu_full_$tag.log
