#!/bin/bash
# v5 bring-up on the GPU box: each shape in its own process (a trapping kernel poisons only that process).
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
export YT8M_WAIT_NOTE=1
for s in "1 64 256 f16" "2 128 256 f16" "2 300 1152 f16" "4 100 256 f16" "7 300 1152 bf16" "3 257 1024 bf16" "37 300 1152 f16" "300 300 1152 f16" "1100 96 256 f16"; do
  echo "=== $s"
  timeout 120 python tools/netvlad_v5_check.py $s 2>&1 | tail -8
done 2>&1 | tee gpurun_out/v5_check.txt
echo "=== timing"
timeout 200 python tools/netvlad_v5_check.py 256 300 1152 f16 time 2>&1 | tail -6 | tee -a gpurun_out/v5_check.txt
echo "=== scan"
timeout 200 python tools/netvlad_v5_scan.py 2>&1 | tail -18 | tee gpurun_out/v6_scan.txt
echo "=== timeline"
timeout 120 python tools/netvlad_v5_timeline.py 256 2>&1 | tail -150 > gpurun_out/v5_timeline.txt
head -70 gpurun_out/v5_timeline.txt
