#!/bin/bash
# ncu evidence for profiles/: launch list of the bench step + full-set capture of the hot kernels.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --train-steps 0 > gpurun_out/launches_run.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"netvlad|gemm_tcgen05|l2norm_rows" -s 5 -c 4 -f -o gpurun_out/prof \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --train-steps 0 > gpurun_out/prof_run.log 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2>gpurun_out/bench.err
cat gpurun_out/bench.json
