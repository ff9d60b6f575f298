#!/bin/bash
# K = 128 bring-up of the one-pass NetVLAD kernel on the GPU box: each shape in its own process.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
export YT8M_WAIT_NOTE=1
for s in "1 64 256 f16" "2 128 256 f16" "2 300 1152 f16" "7 300 1152 bf16" "3 257 1024 bf16" "37 300 1152 f16" "300 300 1152 f16"; do
  echo "=== $s K=128"
  timeout 120 python tools/netvlad_v5_check.py $s notime 128 2>&1 | tail -8
done 2>&1 | tee gpurun_out/k128_check.txt
echo "=== timing"
timeout 200 python tools/netvlad_v5_check.py 512 300 1152 f16 time 128 2>&1 | tail -6 | tee -a gpurun_out/k128_check.txt
timeout 600 python -m pytest tests/test_gpu_edge.py -x -q -k "tiled" 2>&1 | tail -5 | tee -a gpurun_out/k128_check.txt
BENCH_CONFIGS=4 bash tools/r02_quick.sh 2>&1 | tail -5
