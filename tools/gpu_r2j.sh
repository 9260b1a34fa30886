#!/usr/bin/env bash
# the driver's sequence on the current code: full -m gpu suite, smoke, default bench (twice: flakiness check)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$? $(tail -n 1 gpurun_out/pytest_gpu.log)"
grep -E "^(FAILED|ERROR|E  )" gpurun_out/pytest_gpu.log | cut -c1-220 | head
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for rep in 1 2; do
  timeout 420 python bench.py > gpurun_out/bench_default_$rep.log 2> gpurun_out/bench_default_$rep.err; echo "bench #$rep rc=$?"
  grep "\[bench" gpurun_out/bench_default_$rep.err | tail -2
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_default_$rep.log').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}, 'e2e', round(d['e2e']['value']), 'api', round(d['via_reference_api']['value']), 'train', [round(c['value']) for c in d['train']['configs']])
except Exception as e: print('no bench line:', e)
PY
done
