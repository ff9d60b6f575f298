#!/bin/bash
# Round-2 visit 1: GPU test-suite (incl. the BASELINE-shape parity tests), smoke, bench for configs 2-5.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
echo "=== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=8 2>&1 | tail -45 | tee gpurun_out/r02_pytest_gpu.log
echo "=== smoke"
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -5 | tee gpurun_out/r02_smoke.log
for c in 2 3 4 5; do
  echo "=== bench config $c"
  extra=""
  [ "$c" = "5" ] && extra="--batch-sweep"
  timeout 900 python bench.py --config $c --steps 20 --warmup 5 $extra 2>gpurun_out/r02_bench_c$c.err | tee gpurun_out/r02_bench_c$c.json
  tail -3 gpurun_out/r02_bench_c$c.err
done
echo "=== reference arm (config 2)"
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 | tee gpurun_out/r02_bench_ref.json
