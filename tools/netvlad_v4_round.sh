#!/bin/bash
# v4 bring-up on the GPU box: each shape in its own process (a trapping kernel poisons only that process).
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
for s in "2 128 128 f16" "2 300 1152 f16" "3 257 384 bf16" "4 100 256 f16" "7 300 1152 bf16" "300 300 1152 f16" "1 65 1024 f16" "1100 96 128 f16" "75 300 1152 f16" "513 129 256 bf16"; do
  echo "=== $s"
  timeout 120 python tools/netvlad_v4_check.py $s 2>&1 | tail -8
done 2>&1 | tee gpurun_out/v4_check.txt
echo "=== timing"
timeout 200 python tools/netvlad_v4_check.py 256 300 1152 f16 time 2>&1 | tail -6 | tee -a gpurun_out/v4_check.txt
