#!/bin/bash
# Run on the GPU box (under gpurun): the whole -m gpu suite, the attention micro-benchmark and one default bench line.
# Usage: bash tools/run_gpu_suite.sh <tag>
set -u
tag=${1:-x}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -n 15 gpurun_out/pytest_$tag.log
timeout 120 python tools/attn_one.py
timeout 600 python bench.py --steps 60 --warmup 12 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"; tail -c 400 gpurun_out/bench_$tag.err
