#!/usr/bin/env bash
# round 2, first GPU pass: new API-path tests, module/full tests, default bench with per-layer lines
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_api_paths.py tests/test_gpu_full.py tests/test_gpu_modules.py -q -m gpu -p no:cacheprovider -s > gpurun_out/pytest_r2a.log 2>&1; echo "pytest rc=$? $(tail -n 1 gpurun_out/pytest_r2a.log)"
grep -E "^(FAILED|E  )|worst parameter" gpurun_out/pytest_r2a.log | cut -c1-260 | head -30
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
SMAAT_BENCH_LAYERS=1 timeout 900 python bench.py > gpurun_out/bench_r2a.log 2>gpurun_out/bench_r2a.err; echo "bench rc=$?"
tail -5 gpurun_out/bench_r2a.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2a.log').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches','clocks')})
print('e2e', d['e2e'])
print('via_api', d['via_reference_api']['value'], d['via_reference_api']['gap_to_value'], 'alt', d['alt_mode'])
print('parity', d['parity'])
print('eager', d['gpu_eager_baseline'])
print('roofline', {k:d['roofline'][k] for k in ('kernel','bound','frac','frac_hbm','frac_tensor','ms_per_step')})
print('depthwise_roofline', {k:d['depthwise_roofline'][k] for k in ('kernel','achieved','frac','ms_per_step')})
print('cpu_baseline', d['cpu_baseline'])
for k,v in d['kernels'].items(): print(f"   {k:26s} n={v['launches_per_step']:3d} {v['ms_per_step']:7.3f} ms ({100*v['frac_hbm']:5.1f}% hbm) {v['tflops']:6.1f} TF")
PY
grep "^#" gpurun_out/bench_r2a.err | head -70
