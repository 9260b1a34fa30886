#!/usr/bin/env bash
# lean MMA issue loop under elect.sync: parity, layer times per flag set
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for fl in 0 64; do
  SMAAT_DT_FLAGS=$fl timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -p no:cacheprovider -k "dsconv" > gpurun_out/pytest_r2r_$fl.log 2>&1
  rc=$?; echo "flags=$fl dsconv parity rc=$rc $(tail -n 1 gpurun_out/pytest_r2r_$fl.log)"
  if [ $rc -ne 0 ]; then grep -E "^(FAILED|ERROR|E  )" gpurun_out/pytest_r2r_$fl.log | cut -c1-200 | head -5; [ $fl -eq 0 ] && exit 1; fi
done
for fl in 0 1 64 65; do
  echo "== SMAAT_DT_FLAGS=$fl"
  SMAAT_DT_FLAGS=$fl timeout 120 python tools/time_ds.py tf32x3 tmem 2>&1 | awk 'NF>6 {printf "%s ", $(NF-5)} /^sum/ {print $0}'
done
echo "== tf32 flags 65"; SMAAT_DT_FLAGS=65 timeout 120 python tools/time_ds.py tf32 tmem 2>&1 | awk 'NF>6 {printf "%s ", $(NF-5)} /^sum/ {print $0}'
timeout 600 python -m pytest tests/test_gpu_train.py -x -q -m gpu -p no:cacheprovider -k "recompute or upsample or session" > gpurun_out/pytest_r2r_rest.log 2>&1; echo "train tests rc=$? $(tail -n 1 gpurun_out/pytest_r2r_rest.log)"
grep -E "^(FAILED|ERROR|E  )" gpurun_out/pytest_r2r_rest.log | cut -c1-220 | head
