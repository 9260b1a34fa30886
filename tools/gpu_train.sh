#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest -q -m gpu -p no:cacheprovider -rf --tb=short tests/test_gpu_train.py > gpurun_out/train_tests.log 2>&1
echo "train tests rc=$? $(tail -n 1 gpurun_out/train_tests.log)"
grep -E "^(FAILED|E  )" gpurun_out/train_tests.log | cut -c1-260 | head -60
timeout 300 python -m pytest -q -m gpu -p no:cacheprovider tests/test_gpu_kernels.py -k "upsample or pw1x1_stats or dsconv" 2>&1 | tail -2
