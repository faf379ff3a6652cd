#!/bin/bash
# Run on the GPU box (under gpurun): the whole -m gpu suite, then one bench line per workload.
# Usage: bash tools/run_gpu_suite.sh <tag> [extra workloads...]
set -u
tag=${1:-x}; shift
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -n 6 gpurun_out/pytest_$tag.log
timeout 600 python bench.py --steps 60 --warmup 12 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_$tag.err
for wl in "$@"; do
  timeout 600 python bench.py --steps 40 --warmup 10 --workload $wl > gpurun_out/bench_${tag}_$wl.json 2> gpurun_out/bench_${tag}_$wl.err; echo "bench $wl rc=$?"; tail -c 300 gpurun_out/bench_${tag}_$wl.err
done
