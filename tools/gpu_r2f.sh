#!/usr/bin/env bash
# A/B timing of the fused DS layers (both kernels, prefetch on/off) + ncu evidence of the current code
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PATH=$PATH:/usr/local/cuda/bin
timeout 200 python -m pytest tests/test_gpu_train.py tests/test_gpu_api_paths.py -q -m gpu -p no:cacheprovider -k "adam or reference_order_and_serving" 2>&1 | tail -2
echo "== tmem kernel"; timeout 120 python tools/time_ds.py tf32x3 tmem 2>&1 | tail -14
echo "== tmem kernel, no L2 prefetch"; SMAAT_DT_FLAGS=1 timeout 120 python tools/time_ds.py tf32x3 tmem 2>&1 | tail -14
echo "== smem kernel (round 1)"; timeout 120 python tools/time_ds.py tf32x3 smem 2>&1 | tail -14
echo "== tmem kernel, tf32"; timeout 120 python tools/time_ds.py tf32 tmem 2>&1 | tail -14
bash tools/gpu_ncu_r2.sh r02a 2>&1 | tail -12
