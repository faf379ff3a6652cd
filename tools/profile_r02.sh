#!/bin/bash
# Round-2 profiles, run on the GPU box (under gpurun).  Each ncu command runs only after the same command line exited 0 without ncu.
#  (1) launch list (gpu__time_duration per launch) of ONE eager navigation step, DUET cfg-2 and HAMT cfg-3 (tools/step_once.py brackets
#      the step with cudaProfilerStart / Stop);
#  (2) `--set full` capture of the same DUET step, exported to CSV (the .ncu-rep is too large to bring back);
#  (3) launch list of one fine-tuning iteration (cfg-4).
# Usage: bash tools/profile_r02.sh <tag>
set -u
tag=${1:-r02}
mkdir -p gpurun_out
for wl in duet_cfg2 hamt_cfg3; do
  CMD="python tools/step_once.py $wl"
  $CMD > gpurun_out/plain_${tag}_$wl.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/${tag}_${wl}_step_launches.csv $CMD > gpurun_out/ncu_list_${tag}_$wl.log 2>&1
  echo "$wl list rc=$?"
done
CMD2="python tools/step_once.py duet_cfg2"
ncu --set full --clock-control none --import-source on --profile-from-start off -c 120 -f -o gpurun_out/step_$tag $CMD2 > gpurun_out/ncu_full_$tag.log 2>&1
if [ -f gpurun_out/step_$tag.ncu-rep ]; then
  ncu -i gpurun_out/step_$tag.ncu-rep --page raw --csv > gpurun_out/step_${tag}_raw.csv 2> /dev/null
  rm gpurun_out/step_$tag.ncu-rep
fi
CMD3="python bench.py --workload duet_cfg4_train --steps 2 --warmup 1 --no-cpu-baseline"
$CMD3 > gpurun_out/plain3_$tag.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/${tag}_train_launches.csv $CMD3 > gpurun_out/ncu_train_$tag.log 2>&1
echo "train list rc=$?"; wc -l gpurun_out/${tag}_*launches.csv
