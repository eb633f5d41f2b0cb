#!/usr/bin/env bash
# gpurun --gpus 2 -- 'bash tools/two_gpu_check.sh': the data-parallel equivalence check over NCCL and the N = 2 bench line
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 python -m pytest tests/test_parity_headline_gpu.py -m gpu -q -k two_ranks > gpurun_out/r2_two_ranks_nccl.log 2>&1; tail -3 gpurun_out/r2_two_ranks_nccl.log
timeout 300 $TR --master-port 29511 bench.py --gpus 2 --selfcheck > gpurun_out/r2_selfcheck_n2.json 2> gpurun_out/r2_selfcheck_n2.err; tail -c 1500 gpurun_out/r2_selfcheck_n2.json; tail -5 gpurun_out/r2_selfcheck_n2.err | cut -c1-300
timeout 300 $TR --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; tail -1 gpurun_out/r2_bench_n2.err; cut -c1-300 gpurun_out/r2_bench_n2.json
