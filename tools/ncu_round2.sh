#!/usr/bin/env bash
# Round-2 ncu captures for profiles/ (each right after the same command has exited 0 without ncu; ONE gpurun call):
#  (1) --set full of the kernels rewritten this round: two-SM GEMM on the hot shape, fused attention fwd (two-tile) + bwd, TMA-staged LN backward
#  (2) launch list of exactly one timed bench step (bench.py brackets it with cudaProfilerStart/Stop)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cap() {  # name regex cmd...
  local name=$1 rx=$2; shift 2
  timeout 200 "$@" > gpurun_out/plain_$name.log 2>&1 || { echo "plain run of $name failed"; tail -3 gpurun_out/plain_$name.log; return 1; }
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$rx -s 2 -c 1 -f -o gpurun_out/r2_$name "$@" > gpurun_out/ncu_$name.log 2>&1
  tail -1 gpurun_out/ncu_$name.log
}
cap gemm_tc_2sm_48512x1024x1024 gemm_tc python tools/one_gemm.py 48512 1024 1024 1
cap gemm_tc_2sm_48512x4096x1024_gelu gemm_tc python tools/one_gemm.py 48512 4096 1024 1 0 0 1
cap fmha_fwd2 fmha_fwd2 python tools/one_fmha.py 32 16 1516 fwd
cap fmha_bwd fmha_bwd_kernel python tools/one_fmha.py 32 16 1516 both
cap ln_bwd_tma ln_bwd_tma python tools/one_hbm_kernels.py ln
timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_plain_r2.log 2>&1 || { echo "plain bench failed"; tail -3 gpurun_out/bench_plain_r2.log; exit 1; }
tail -1 gpurun_out/bench_plain_r2.log | cut -c1-200
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_r2.log 2>&1
wc -l gpurun_out/launches_r2.csv
