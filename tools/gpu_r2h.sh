#!/usr/bin/env bash
# full -m gpu suite exactly as the driver runs it + smoke + default bench, then the ncu evidence of the same code
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$? $(tail -n 1 gpurun_out/pytest_gpu.log)"
grep -E "^(FAILED|ERROR|E  )" gpurun_out/pytest_gpu.log | cut -c1-220 | head
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench_default.log 2>gpurun_out/bench_default.err; echo "bench rc=$?"
tail -2 gpurun_out/bench_default.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_default.log').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches','clocks')})
print('e2e', d['e2e']['value'], 'via_api', d['via_reference_api']['value'], d['via_reference_api']['gap_to_value'], 'alt', d['alt_mode'])
print('parity', d['parity']); print('eager', d['gpu_eager_baseline'])
print('roofline', {k:d['roofline'][k] for k in ('kernel','bound','frac','frac_hbm','frac_tensor','ms_per_step','traffic')})
print('depthwise_roofline', {k:d['depthwise_roofline'][k] for k in ('achieved','frac','ms_per_step')})
print('train', json.dumps(d['train'])[:800]); print('cpu', d['cpu_baseline'])
for k,v in d['kernels'].items(): print(f"   {k:28s} n={v['launches_per_step']:3d} {v['ms_per_step']:7.3f} ms ({100*v['frac_hbm']:5.1f}% hbm) {v['tflops']:6.1f} TF")
PY
bash tools/gpu_ncu_r2.sh r02 2>&1 | tail -8
du -sh gpurun_out
