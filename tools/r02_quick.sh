#!/bin/bash
# quick GPU check after a GEMM change: kernel + model tests, then bench config 2
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py tests/test_gpu_config_shapes.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -6
for c in ${BENCH_CONFIGS:-2}; do
  timeout 600 python bench.py --config $c --steps 20 --warmup 5 --no-cpu-baseline ${BENCH_EXTRA} 2>gpurun_out/q_bench_c$c.err | tee gpurun_out/q_bench_c$c.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('config', d['config']['baseline_config'], {k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], 'train', (d.get('train_step') or {}).get('ms_per_step'))
print('roofline', {k:r.get(k) for k in ('kernel_ms','achieved','frac')})"
  tail -2 gpurun_out/q_bench_c$c.err
done
