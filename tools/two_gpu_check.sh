set -u
cd $GRAFT_REPO_ROOT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_parity_headline_gpu.py -m gpu -q -k two_ranks 2>&1 | tail -4
timeout 600 $TR --master-port 29511 bench.py --gpus 2 --selfcheck 2>&1 | tail -3 | cut -c1-1500
timeout 600 $TR --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 2>&1 | tail -2 | cut -c1-420
TSW_DDP_BF16=1 timeout 600 $TR --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 2>&1 | tail -2 | cut -c1-420
