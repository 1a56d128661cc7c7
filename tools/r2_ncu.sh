#!/bin/bash
# Developer tool (GPU box): ncu --set full captures of one decode launch per corpus class.  Usage: tools/r2_ncu.sh "T M" lz10 [lib variant]
CLASSES=${1:-"M"}
FMT=${2:-lz10}
VAR=$3
mkdir -p gpurun_out
if [ -n "$VAR" ]; then export AURORA_CUDA_LIB=$PWD/auroralib/compression_b200/variants/libaurora_cuda_$VAR.so; fi
for c in $CLASSES; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_ -s 1 -c 1 -f -o gpurun_out/prof_${FMT}_${c}${VAR:+_$VAR} \
     python tools/perf_probe.py --formats $FMT --classes $c --streams 5032 --iters 1 > gpurun_out/ncu_${FMT}_${c}.log 2>&1
  tail -2 gpurun_out/ncu_${FMT}_${c}.log
done
