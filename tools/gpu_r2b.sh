#!/usr/bin/env bash
# round 2: first run of the TMEM-operand fused kernel -- parity first (short timeouts: a hang must not eat the box), then timing
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -k "dsconv" -x > gpurun_out/pytest_r2b_ds.log 2>&1; rc=$?
echo "dsconv kernel tests rc=$rc $(tail -n 1 gpurun_out/pytest_r2b_ds.log)"
grep -E "^(FAILED|E  )" gpurun_out/pytest_r2b_ds.log | cut -c1-240 | head -20
if [ $rc -ne 0 ]; then
  timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -k "dsconv" > gpurun_out/pytest_r2b_ds_all.log 2>&1
  echo "all: $(tail -n 1 gpurun_out/pytest_r2b_ds_all.log)"; grep -E "^FAILED" gpurun_out/pytest_r2b_ds_all.log | cut -c1-200 | head -60
  exit 0
fi
timeout 900 python -m pytest tests/test_gpu_api_paths.py tests/test_gpu_full.py tests/test_gpu_modules.py -q -m gpu -p no:cacheprovider > gpurun_out/pytest_r2b.log 2>&1; echo "pytest rc=$? $(tail -n 1 gpurun_out/pytest_r2b.log)"
grep -E "^(FAILED|E  )" gpurun_out/pytest_r2b.log | cut -c1-240 | head -20
for impl in 1 0; do
  SMAAT_DS_IMPL=$impl SMAAT_BENCH_LAYERS=1 timeout 600 python bench.py --no-cpu-baseline --no-alt > gpurun_out/bench_r2b_impl$impl.log 2>gpurun_out/bench_r2b_impl$impl.err; echo "bench impl=$impl rc=$?"
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_r2b_impl$impl.log').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','clocks')})
print('roofline', {k:d['roofline'][k] for k in ('kernel','bound','frac','frac_hbm','frac_tensor','ms_per_step')})
PY
  grep "^# smaat_dsconv" gpurun_out/bench_r2b_impl$impl.err | sort -u
done
