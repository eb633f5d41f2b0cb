#!/usr/bin/env bash
# ncu captures for profiles/ (each after the same command has exited 0 without ncu):
#  (1) --set full of the HBM-bound kernels rewritten this round, (2) launch list of exactly one timed bench step (cudaProfilerStart/Stop around it)
set -u
mkdir -p gpurun_out
timeout 120 python tools/one_hbm_kernels.py all > gpurun_out/one_hbm_plain.log 2>&1 || { echo "plain run failed"; tail -3 gpurun_out/one_hbm_plain.log; exit 1; }
for k in decode_attention ln_bwd_fused colsum_partial; do
  arg=decode; [ $k = ln_bwd_fused ] && arg=ln; [ $k = colsum_partial ] && arg=colsum
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/r1_$k python tools/one_hbm_kernels.py $arg > gpurun_out/ncu_$k.log 2>&1
  tail -1 gpurun_out/ncu_$k.log
done
timeout 200 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_plain4.log 2>&1 || { echo "plain bench failed"; exit 1; }
tail -1 gpurun_out/bench_plain4.log | cut -c1-200
timeout 700 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1_v4.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench4.log 2>&1
wc -l gpurun_out/launches_r1_v4.csv
