#!/bin/bash
# Runs each GPU test group in its own process (a wedged / faulting kernel poisons only its group)
# and a quick per-kernel timing pass.  Output -> gpurun_out/.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/bringup_smi.txt 2>&1
status=0
for k in "l2norm" "linear" "moe or group_max" "lstm" "attn" "netvlad" "xent or topk or context"; do
  tag=$(echo "$k" | tr ' ' '_')
  echo "=== group: $k"
  timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "$k" -p no:cacheprovider 2>&1 | tail -25 | tee "gpurun_out/bringup_${tag}.log"
  rc=${PIPESTATUS[0]}
  [ "$rc" != "0" ] && status=1
done
echo "=== quick bench"
timeout 900 python tools/gpu_quick_bench.py 2>&1 | tail -30 | tee gpurun_out/quick_bench.log
exit $status
