#!/bin/bash
# Developer tool (GPU box): flag-LZ parity tests, then per-class probes of main + variants.  Usage: tools/r2_quick.sh "v1 v2" [formats] [classes]
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_decode_gpu.py tests/test_wrappers.py -x -q > gpurun_out/quick_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/quick_tests.log
tail -4 gpurun_out/quick_tests.log
bash tools/r2_variants.sh "$1" "${2:-lz10}" "${3:-T,M,X,mix}"
