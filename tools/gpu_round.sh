#!/bin/bash
# One GPU visit: full GPU test-suite, smoke, bench, ncu launch list + full capture of the hot kernels.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
echo "=== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -30 | tee gpurun_out/pytest_gpu.log
echo "=== smoke"
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "=== bench"
timeout 600 python bench.py --steps 20 --warmup 5 2>gpurun_out/bench.err | tee gpurun_out/bench.json
tail -5 gpurun_out/bench.err
if [ "$1" != "noprof" ]; then
echo "=== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --train-steps 0 > gpurun_out/launches_run.log 2>&1
tail -3 gpurun_out/launches_run.log
echo "=== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"netvlad|gemm_tcgen05|l2norm_rows" -s 5 -c 4 -f -o gpurun_out/prof \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --train-steps 0 > gpurun_out/prof_run.log 2>&1
tail -3 gpurun_out/prof_run.log
ls -la gpurun_out/
fi
