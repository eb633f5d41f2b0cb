#!/usr/bin/env bash
# gpurun --gpus N -- 'bash tools/n2_nccl_ctas.sh N': data-parallel step with NCCL's CTA count capped (the GEMM work list is dynamic)
set -u
cd "$(dirname "$0")/.."
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
run() {  # tag env...
  local tag=$1; shift
  env "$@" timeout 300 $TR --master-port $((29600 + RANDOM % 200)) bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/nc_${N}_$tag.json 2> gpurun_out/nc_${N}_$tag.err
  echo "$tag: $(tail -1 gpurun_out/nc_${N}_$tag.err | sed 's/.*tcgen05 GEMM/GEMM/')"
}
run default TSW_X=0
run maxctas8 NCCL_MAX_CTAS=8
run maxctas4 NCCL_MAX_CTAS=4
run maxctas16 NCCL_MAX_CTAS=16
run default2 TSW_X=0
