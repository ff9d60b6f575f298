#!/bin/bash
# Round-2 closing evidence: GPU suite, smoke, bench lines of configs 2-5 + the reference arm, ncu launch list, memcheck of the
# training step's new kernels.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
echo "=== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=6 2>&1 | tail -30 | tee gpurun_out/r02g_pytest_gpu.log
echo "=== smoke"
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -4 | tee gpurun_out/r02g_smoke.log
for c in 2 3 4 5; do
  echo "=== bench config $c"
  timeout 900 python bench.py --config $c --steps 20 --warmup 5 2>gpurun_out/r02g_bench_c$c.err | tee gpurun_out/r02g_bench_c$c.json | cut -c1-400
  tail -2 gpurun_out/r02g_bench_c$c.err
done
echo "=== reference arm (config 2)"
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 | tee gpurun_out/r02g_bench_ref.json | cut -c1-300
echo "=== memcheck: training-step kernels"
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_targets.py train 2>&1 | grep -v "^$" | tail -8 | tee gpurun_out/r02g_sanitizer_train.log
echo "=== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 160 --csv --log-file gpurun_out/r02g_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --train-steps 0 > gpurun_out/r02g_launches_run.log 2>&1
tail -1 gpurun_out/r02g_launches_run.log | cut -c1-200
