#!/usr/bin/env bash
# 8-GPU scaling check: eval bench (sharded batch, no collective) and DDP training (global batch 256)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.log 2>gpurun_out/bench_n$N.err
echo "bench N=$N rc=$?"; tail -n 1 gpurun_out/bench_n$N.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','n_gpus','ms_per_step','gpu_launches','clocks')}, 'e2e', d['e2e']['value'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench_train.py --global-batch 256 --steps 8 --warmup 3 > gpurun_out/train_n$N.log 2>&1
echo "ddp train N=$N rc=$?"; tail -n 1 gpurun_out/train_n$N.log | cut -c1-420
