#!/usr/bin/env bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m pytest -q -m gpu -p no:cacheprovider tests/test_gpu_kernels.py -k "upsample or maxpool" 2>&1 | tail -1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --mode tf32x3 > gpurun_out/bench_n$N.log 2>&1
echo "bench N=$N rc=$?"
python - "$N" <<'PY'
import json,sys
n=sys.argv[1]
txt=open(f'gpurun_out/bench_n{n}.log').read().strip().splitlines()
try:
    d=json.loads(txt[-1]); print({k:d[k] for k in ('value','n_gpus','ms_per_step','scaling','gpu_launches','clocks')}, 'e2e', d['e2e']['value'])
    print({k: round(v['ms_per_step'],3) for k,v in d['kernels'].items()})
except Exception as e:
    print('no json', e); print('\n'.join(txt[-15:]))
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | cut -c1-400
