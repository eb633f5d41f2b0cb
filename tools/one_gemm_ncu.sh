#!/usr/bin/env bash
# ncu captures for profiles/: (1) launch list of one bench step, (2) --set full of the dominant kernel on the hot shape
set -u
python tools/one_gemm.py 48512 1024 1024 1 > gpurun_out/one_gemm_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 1 -c 1 -o gpurun_out/r1_gemm_tc_48512x1024x1024 python tools/one_gemm.py 48512 1024 1024 1 > gpurun_out/ncu_gemm.log 2>&1
tail -1 gpurun_out/ncu_gemm.log
python bench.py --model medium --batch 32 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 7600 -c 2600 --csv --log-file gpurun_out/launches_r1_v2.csv python bench.py --model medium --batch 32 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/bench_plain.log | cut -c1-200
wc -l gpurun_out/launches_r1_v2.csv
