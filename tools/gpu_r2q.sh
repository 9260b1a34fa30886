#!/usr/bin/env bash
# lean MMA issue loop (+ optional second issuing warp): parity, layer times, then the training-path changes
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for fl in 0 64; do
  SMAAT_DT_FLAGS=$fl timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -p no:cacheprovider -k "dsconv" > gpurun_out/pytest_r2q_$fl.log 2>&1
  rc=$?; echo "flags=$fl dsconv parity rc=$rc $(tail -n 1 gpurun_out/pytest_r2q_$fl.log)"
  if [ $rc -ne 0 ]; then grep -E "^(FAILED|ERROR|E  )" gpurun_out/pytest_r2q_$fl.log | cut -c1-200 | head -5; [ $fl -eq 0 ] && exit 1; fi
done
for fl in 0 64 65; do
  echo "== SMAAT_DT_FLAGS=$fl"
  SMAAT_DT_FLAGS=$fl timeout 120 python tools/time_ds.py tf32x3 tmem 2>&1 | awk '{printf "%s ", $(NF-5)} END {print ""}'
done
echo "== tf32"; timeout 120 python tools/time_ds.py tf32 tmem 2>&1 | awk '{printf "%s ", $(NF-5)} END {print ""}'
timeout 600 python -m pytest tests/test_gpu_train.py tests/test_gpu_kernels.py -x -q -m gpu -p no:cacheprovider -k "not dsconv" > gpurun_out/pytest_r2q_rest.log 2>&1; echo "train+kernel tests rc=$? $(tail -n 1 gpurun_out/pytest_r2q_rest.log)"
grep -E "^(FAILED|ERROR|E  )" gpurun_out/pytest_r2q_rest.log | cut -c1-220 | head
timeout 300 python tools/train_breakdown.py 32 > gpurun_out/train_breakdown_r02.txt 2>&1; head -30 gpurun_out/train_breakdown_r02.txt
SMAAT_BENCH_LAYERS=1 timeout 420 python bench.py > gpurun_out/bench_r2q.log 2> gpurun_out/bench_r2q.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_r2q.log').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'])
    print('train', json.dumps(d['train'])[:900])
    for k,v in d['kernels'].items(): print(f"   {k:28s} n={v['launches_per_step']:3d} {v['ms_per_step']:7.3f} ms ({100*v['frac_hbm']:5.1f}% hbm) {v['tflops']:6.1f} TF")
except Exception as e: print('no bench line:', e)
PY
grep "^# smaat_dsconv" gpurun_out/bench_r2q.err | sort -u | head -20
