#!/bin/bash
# Developer tool (GPU box): compute-sanitizer memcheck / racecheck / synccheck over tools/sanitize_smoke.py; logs -> gpurun_out/
mkdir -p gpurun_out
for tool in memcheck synccheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_smoke.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool exit $?" >> gpurun_out/sanitize_$tool.log
  echo "== $tool"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|mismatches|exit|hazard" gpurun_out/sanitize_$tool.log | sort | uniq -c | head -12
done
