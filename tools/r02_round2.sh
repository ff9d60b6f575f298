#!/bin/bash
# GPU test-suite + bench config 2 (+4) after a kernel change
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
echo "=== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -15 | tee gpurun_out/r02_pytest_gpu.log
echo "=== smoke"
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3
for c in ${BENCH_CONFIGS:-2}; do
  echo "=== bench config $c"
  timeout 900 python bench.py --config $c --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/r02_bench_c$c.err | tee gpurun_out/r02_bench_c$c.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], 'train', (d.get('train_step') or {}).get('ms_per_step'))
print('roofline', {k:r.get(k) for k in ('kernel_ms','achieved','frac')})"
  tail -3 gpurun_out/r02_bench_c$c.err
done
