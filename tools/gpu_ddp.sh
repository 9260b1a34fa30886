#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench_train.py --global-batch $((64*N)) --steps 5 --warmup 2 > gpurun_out/train_n$N.log 2>&1
echo "ddp train N=$N rc=$?"; tail -n 2 gpurun_out/train_n$N.log | cut -c1-400
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.log 2>&1
echo "bench N=$N rc=$?"; tail -n 1 gpurun_out/bench_n$N.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','n_gpus','ms_per_step','gpu_launches')}, d['e2e']['value'])"
