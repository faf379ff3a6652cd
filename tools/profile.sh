#!/bin/bash
# Run on the GPU box (under gpurun).  (1) launch list of a short bench.py run (every kernel with its device time),
# (2) one --set full capture of every kernel of ONE eager navigation step (tools/step_once.py brackets it with
# cudaProfilerStart/Stop), exported to CSV on the box; the .ncu-rep is kept only while gpurun_out stays small.
# Each ncu command runs only after the same command line exited 0 without ncu.   Usage: bash tools/profile.sh <tag> [workload]
set -u
tag=${1:-r01}
wl=${2:-duet_cfg2}
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload $wl"
$CMD > gpurun_out/plain_$tag.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_$tag.csv $CMD > gpurun_out/ncu_list_$tag.log 2>&1
CMD2="python tools/step_once.py $wl"
$CMD2 > gpurun_out/plain2_$tag.log 2>&1 &&
ncu --set full --clock-control none --import-source on --profile-from-start off -c 120 -f -o gpurun_out/step_$tag $CMD2 > gpurun_out/ncu_full_$tag.log 2>&1
if [ -f gpurun_out/step_$tag.ncu-rep ]; then
  ncu -i gpurun_out/step_$tag.ncu-rep --page raw --csv > gpurun_out/step_${tag}_raw.csv 2> /dev/null
  ls -la gpurun_out/step_$tag.ncu-rep
  # keep the report only if it fits the 64 MiB return budget comfortably
  sz=$(stat -c %s gpurun_out/step_$tag.ncu-rep)
  if [ "$sz" -gt 45000000 ]; then rm gpurun_out/step_$tag.ncu-rep; echo "report dropped (too large), CSV kept"; fi
fi
tail -n 2 gpurun_out/ncu_list_$tag.log | cut -c1-300; tail -n 3 gpurun_out/ncu_full_$tag.log
