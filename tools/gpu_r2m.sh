#!/usr/bin/env bash
# final validation of the round: full -m gpu suite, smoke, default bench, compute-sanitizer memcheck over the kernel / module tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PATH=$PATH:/usr/local/cuda/bin
timeout 900 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$? $(tail -n 1 gpurun_out/pytest_gpu.log)"
grep -E "^(FAILED|ERROR|E  )" gpurun_out/pytest_gpu.log | cut -c1-220 | head
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
SMAAT_BENCH_LAYERS=1 timeout 420 python bench.py > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_default.log').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches','clocks')})
    print('e2e', d['e2e']['value'], 'via_api', d['via_reference_api']['value'], d['via_reference_api']['gap_to_value'], 'alt', d['alt_mode'])
    print('parity', d['parity']); print('eager', d['gpu_eager_baseline'])
    print('roofline', {k:d['roofline'][k] for k in ('kernel','bound','frac','frac_hbm','frac_tensor','ms_per_step','traffic')})
    print('depthwise_roofline', {k:d['depthwise_roofline'][k] for k in ('achieved','frac','ms_per_step','traffic')})
    print('train', json.dumps(d['train'])[:600]); print('cpu', d['cpu_baseline'])
    for k,v in d['kernels'].items(): print(f"   {k:28s} n={v['launches_per_step']:3d} {v['ms_per_step']:7.3f} ms ({100*v['frac_hbm']:5.1f}% hbm) {v['tflops']:6.1f} TF")
except Exception as e: print('no bench line:', e)
PY
grep "^# smaat" gpurun_out/bench_default.err | sort -u | head -40
{
echo "# compute-sanitizer memcheck on the round's final code (B200, under gpurun)"
echo "## memcheck: tests/test_gpu_kernels.py (all kernels incl. both fused DS-conv kernels, CBAM, upsample, transposed conv)"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -x 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|out of bounds" | head -8
echo "## memcheck: tests/test_gpu_modules.py (every eval-mode module case in the three pointwise modes)"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_modules.py -q -m gpu -p no:cacheprovider -x 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|out of bounds" | head -8
} > gpurun_out/compute_sanitizer_r02.txt 2>&1
cat gpurun_out/compute_sanitizer_r02.txt
