#!/bin/bash
# closing check after the last kernel changes: full GPU suite, smoke, bench lines of configs 2 and 3
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -4 | tee gpurun_out/r02h_pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2 | tee gpurun_out/r02h_smoke.log
for c in 2 3; do
  timeout 900 python bench.py --config $c --steps 20 --warmup 5 2>/dev/null | tee gpurun_out/r02h_bench_c$c.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().splitlines()[-1]); t=d['train_step']; r=d['roofline']
print('config $c', round(d['value']), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), d['e2e'].get('h2d_gb_per_s_per_gpu'), 'train', round(t['ms_per_step'],3), 'frac', round(r['frac'],3), d['clocks'])"
done
