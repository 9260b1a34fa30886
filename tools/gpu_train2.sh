#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest -q -m gpu -p no:cacheprovider -x --tb=short tests/test_gpu_train.py -k "train_session or training_step" 2>&1 | tail -12
for g in "" "--no-graph"; do
  timeout 600 python bench_train.py --steps 8 --warmup 3 $g 2>&1 | tail -1
done
timeout 600 python bench_train.py --steps 8 --warmup 3 --mode tf32 2>&1 | tail -1
