#!/usr/bin/env bash
# round 2: everything written while the GPU pool was busy, with failure isolation: new kernels first (short timeouts), then
# the whole -m gpu suite, stage timers, and the default bench with per-layer lines
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { # name, timeout, args...
  local name=$1 to=$2; shift 2
  timeout $to python -m pytest "$@" -q -m gpu -p no:cacheprovider > gpurun_out/pytest_$name.log 2>&1; local rc=$?
  echo "[$name] rc=$rc $(tail -n 1 gpurun_out/pytest_$name.log)"
  grep -E "^(FAILED|ERROR|E  )" gpurun_out/pytest_$name.log | cut -c1-220 | head -12
  return $rc
}
run ds 300 tests/test_gpu_kernels.py -k "dsconv"
run cbam_up 300 tests/test_gpu_kernels.py -k "cbam or upsample or pw1x1 or pointwise"
run train 600 tests/test_gpu_train.py
run api 600 tests/test_gpu_api_paths.py
run rest 900 tests/test_gpu_modules.py tests/test_gpu_full.py tests/test_metrics.py tests/test_data.py tests/test_gpu_kernels.py -k "not dsconv and not cbam and not upsample and not pw1x1 and not pointwise"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for a in "12 288 64" "64 288 64" "128 288 64" "64 144 128" "256 144 128" "256 72 256"; do timeout 120 python tools/dt_timing.py $a tf32x3 2>&1 | tail -5; done
SMAAT_BENCH_LAYERS=1 timeout 900 python bench.py > gpurun_out/bench_r2e.log 2>gpurun_out/bench_r2e.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_r2e.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2e.log').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches','clocks')})
print('e2e', d['e2e']['value'], 'via_api', d['via_reference_api']['value'], d['via_reference_api']['gap_to_value'], 'alt', d['alt_mode'])
print('parity', d['parity']); print('eager', d['gpu_eager_baseline'])
print('roofline', {k:d['roofline'][k] for k in ('kernel','bound','frac','frac_hbm','frac_tensor','ms_per_step')})
print('depthwise_roofline', {k:d['depthwise_roofline'][k] for k in ('achieved','frac','ms_per_step')})
print('train', json.dumps(d['train'])[:1500])
for k,v in d['kernels'].items(): print(f"   {k:28s} n={v['launches_per_step']:3d} {v['ms_per_step']:7.3f} ms ({100*v['frac_hbm']:5.1f}% hbm) {v['tflops']:6.1f} TF")
PY
grep "^#" gpurun_out/bench_r2e.err | sort -u | head -60
SMAAT_PW_ATMEM=0 SMAAT_BENCH_LAYERS=1 timeout 600 python bench.py --no-cpu-baseline --no-alt --no-train > gpurun_out/bench_r2e_noatm.log 2>gpurun_out/bench_r2e_noatm.err; echo "bench(no ATM) rc=$?"
grep "^# smaat_pw1x1" gpurun_out/bench_r2e_noatm.err | sort -u
