#!/bin/bash
# Developer tool (GPU box): GPU parity suite, the full bench line (all configs), reference arm.  Usage: tools/r2_full.sh [extra bench args]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/full_gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/full_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/full_tests.log
tail -4 gpurun_out/full_tests.log
timeout 900 python bench.py "$@" > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
echo "bench exit $?"
tail -c 600 gpurun_out/bench_full.err
python tools/bench_summary.py gpurun_out/bench_full.json 2>&1 | tail -20
