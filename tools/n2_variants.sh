#!/usr/bin/env bash
# gpurun --gpus N -- 'bash tools/n2_variants.sh N': data-parallel step time under different gradient-exchange settings (same box)
set -u
cd "$(dirname "$0")/.."
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
run() {  # tag env...
  local tag=$1; shift
  env "$@" timeout 300 $TR --master-port $((29600 + RANDOM % 200)) bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/nv_${N}_$tag.json 2> gpurun_out/nv_${N}_$tag.err
  echo "$tag: $(tail -1 gpurun_out/nv_${N}_$tag.err | sed 's/.*clean step/clean step/')"
}
run fp32 TSW_DDP_BF16=0
run bf16 TSW_DDP_BF16=1
run fp32_minctas16 TSW_DDP_BF16=0 NCCL_MIN_CTAS=16
run bf16_bucket256 TSW_DDP_BF16=1 TSW_DDP_BUCKET_MB=256
run bf16_reserve8 TSW_DDP_BF16=1 TSW_SM_RESERVE=8
NCCL_DEBUG=INFO TSW_DDP_BF16=0 timeout 300 $TR --master-port 29999 bench.py --gpus $N --steps 2 --warmup 3 2>&1 | grep -E "NCCL INFO.*(Channel|channels|Connected|Algo|NVLS|CTA|nThreads|comm )" | head -20 > gpurun_out/nv_${N}_nccl_info.txt
head -12 gpurun_out/nv_${N}_nccl_info.txt | cut -c1-220
