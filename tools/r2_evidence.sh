#!/bin/bash
# Developer tool (GPU box): the profiling artefacts of the round: per-class ncu captures, the bench launch list, one full capture of the headline launch
mkdir -p gpurun_out
bash tools/r2_ncu.sh "T M X" lz10
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:decode_|encode_|size_order|lcg_|is_match|reverse_' -c 400 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 5 --warmup 3 --no-configs --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
echo "launch list exit $?"; grep -c decode_flaglz gpurun_out/launches_bench.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_flaglz -s 3 -c 1 -f -o gpurun_out/prof_lz10_c2_full python bench.py --steps 2 --warmup 3 --no-configs --no-cpu-baseline > gpurun_out/bench_under_ncu_full.log 2>&1
echo "full capture exit $?"
