#!/usr/bin/env bash
# three producer groups with per-(stage, group) hand-back barriers: parity first (stop at the first failure), then layer times
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for fl in 0 64; do
  SMAAT_DT_FLAGS=$fl timeout 150 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -p no:cacheprovider -k "dsconv" > gpurun_out/pytest_r2w_$fl.log 2>&1
  rc=$?; echo "flags=$fl dsconv parity rc=$rc $(tail -n 1 gpurun_out/pytest_r2w_$fl.log)"
  if [ $rc -ne 0 ]; then grep -E "^(FAILED|ERROR|E  )" gpurun_out/pytest_r2w_$fl.log | cut -c1-200 | head -5; exit 1; fi
done
for fl in 0 64; do
  echo "== SMAAT_DT_FLAGS=$fl"
  SMAAT_DT_FLAGS=$fl timeout 60 python tools/time_ds.py tf32x3 tmem 2>&1 | awk 'NF>6 {printf "%s ", $(NF-5)} /^sum/ {print $0}'
  [ ${PIPESTATUS[0]} -eq 124 ] && { echo "TIMEOUT"; exit 1; }
done
echo "== tf32"; timeout 60 python tools/time_ds.py tf32 tmem 2>&1 | awk 'NF>6 {printf "%s ", $(NF-5)} /^sum/ {print $0}'
