#!/bin/bash
# Developer tool (GPU box): per-class probes of built variants without the test suite.  Usage: tools/r2_variants.sh "v1 v2" [formats] [classes]
VARS=${1:-""}; FMTS=${2:-lz10}; CLS=${3:-T,M,X,mix}
mkdir -p gpurun_out
: > gpurun_out/variants_perf.log
for v in main $VARS; do
  echo "== $v" >> gpurun_out/variants_perf.log
  if [ "$v" = main ]; then unset AURORA_CUDA_LIB; else export AURORA_CUDA_LIB=$PWD/auroralib/compression_b200/variants/libaurora_cuda_$v.so; fi
  EXTRA=""; if [ "$v" = noreplay ]; then EXTRA="--no-verify"; fi
  timeout 600 python tools/perf_probe.py --formats $FMTS --classes $CLS --streams 16384 $EXTRA >> gpurun_out/variants_perf.log 2>&1
done
cat gpurun_out/variants_perf.log
