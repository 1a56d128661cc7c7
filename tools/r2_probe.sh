#!/bin/bash
# Developer tool (GPU box): parity tests + per-class LZ10 probes of the built variants.  Usage: tools/r2_probe.sh "var1 var2 ..." [formats]
VARS=${1:-""}
FMTS=${2:-lz10}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/probe_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/probe_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/probe_tests.log
tail -3 gpurun_out/probe_tests.log
echo "== main" > gpurun_out/probe_perf.log
timeout 600 python tools/perf_probe.py --formats $FMTS --streams 16384 >> gpurun_out/probe_perf.log 2>&1
for v in $VARS; do
  echo "== $v" >> gpurun_out/probe_perf.log
  EXTRA=""
  if [ "$v" = noreplay ]; then EXTRA="--no-verify"; fi
  AURORA_CUDA_LIB=$PWD/auroralib/compression_b200/variants/libaurora_cuda_$v.so timeout 600 python tools/perf_probe.py --formats $FMTS --streams 16384 $EXTRA >> gpurun_out/probe_perf.log 2>&1
done
cat gpurun_out/probe_perf.log
