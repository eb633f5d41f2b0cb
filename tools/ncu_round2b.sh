#!/usr/bin/env bash
# Round-2 (second half) ncu captures for profiles/: the TMA-store epilogue kinds of the tcgen05 GEMM, the single-query-tile attention
# backward, and the launch list of one timed step.  Each capture runs right after the same command exited 0 without ncu.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cap() {  # name regex cmd...
  local name=$1 rx=$2; shift 2
  timeout 200 "$@" > gpurun_out/plain_$name.log 2>&1 || { echo "plain run of $name failed"; tail -3 gpurun_out/plain_$name.log; return 1; }
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$rx -s 2 -c 1 -f -o gpurun_out/r2b_$name "$@" > gpurun_out/ncu_$name.log 2>&1
  tail -1 gpurun_out/ncu_$name.log
}
cap gemm_tma_48512x1024x1024_bias_res gemm_tc python tools/one_gemm.py 48512 1024 1024 1 0 0 0 res
cap gemm_tma_48512x4096x1024_gelu_grad gemm_tc python tools/one_gemm.py 48512 4096 1024 1 0 0 3
cap gemm_tma_48512x4096x1024_mulaux_colsum gemm_tc python tools/one_gemm.py 48512 4096 1024 1 0 1 4
cap fmha_bwd_q1 fmha_bwd_q1 python tools/one_fmha_q1.py
timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_plain_r2b.log 2>&1 || { echo "plain bench failed"; tail -3 gpurun_out/bench_plain_r2b.log; exit 1; }
tail -1 gpurun_out/bench_plain_r2b.log | cut -c1-200
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2b.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_r2b.log 2>&1
wc -l gpurun_out/launches_r2b.csv
