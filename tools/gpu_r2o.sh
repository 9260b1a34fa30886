#!/usr/bin/env bash
# pipelined epilogue + dual MMA issuers: parity, then the knock-out matrix (SMAAT_DT_FLAGS) that shows the binding stage
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for fl in 0 64; do
  SMAAT_DT_FLAGS=$fl timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -p no:cacheprovider -k "dsconv" > gpurun_out/pytest_r2o_$fl.log 2>&1
  rc=$?; echo "flags=$fl dsconv parity rc=$rc $(tail -n 1 gpurun_out/pytest_r2o_$fl.log)"
  if [ $rc -ne 0 ]; then grep -E "^(FAILED|ERROR|E  )" gpurun_out/pytest_r2o_$fl.log | cut -c1-200 | head -5; [ $fl -eq 0 ] && exit 1; fi
done
for fl in 0 64 2 4 8 16 32 66 72 80; do
  echo "== SMAAT_DT_FLAGS=$fl"
  SMAAT_DT_FLAGS=$fl timeout 120 python tools/time_ds.py tf32x3 tmem 2>&1 | awk '{printf "%s %s %s %s | ", $2, $4, $6, $7} END {print ""}'
done
