#!/usr/bin/env bash
# gpurun --gpus N -- 'bash tools/n2_sched_ab.sh N': data-parallel step, alternating runs on one box: dynamic (cluster launch control) vs static
# work lists.  GEMM: dynamic by default, TSW_GEMM_STATIC=1 switches it off; attention backward: static by default, TSW_FMHA_DYNAMIC=1 on.
set -u
cd "$(dirname "$0")/.."
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
run() {  # tag env...
  local tag=$1; shift
  env "$@" timeout 300 $TR --master-port $((29600 + RANDOM % 200)) bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/ab_${N}_$tag.json 2> gpurun_out/ab_${N}_$tag.err
  echo "$tag: $(tail -1 gpurun_out/ab_${N}_$tag.err | sed 's/.*tcgen05 GEMM/GEMM/')"
}
run warm TSW_X=0
run gemm_dynamic_1 TSW_X=0
run gemm_static_1 TSW_GEMM_STATIC=1
run gemm_dynamic_2 TSW_X=0
run gemm_static_2 TSW_GEMM_STATIC=1
run fmha_dynamic_1 TSW_FMHA_DYNAMIC=1
run fmha_static_1 TSW_X=0
