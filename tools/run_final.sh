#!/bin/bash
# Round-end validation on the GPU box: smoke(), the whole -m gpu suite, the default bench line, the fine-tuning bench line.
# Usage: bash tools/run_final.sh <tag>
set -u
tag=${1:-final}
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/pytest_$tag.log
timeout 600 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_$tag.err
timeout 600 python bench.py --workload duet_cfg4_train --steps 6 --warmup 3 > gpurun_out/train_$tag.json 2> gpurun_out/train_$tag.err; echo "train rc=$?"; tail -c 300 gpurun_out/train_$tag.err
