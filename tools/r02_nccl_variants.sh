#!/bin/bash
# train-step time of BASELINE config 2 at N ranks under a few NCCL settings (how the gradient pieces share the GPU with the backward)
cd "$(dirname "$0")/.." || exit 1
N=${1:-2}
mkdir -p gpurun_out
for v in "NCCL_DEBUG=WARN" "NCCL_PROTO=Simple" "NCCL_MAX_CTAS=16" "NCCL_PROTO=Simple NCCL_MAX_CTAS=16"; do
  tag=$(echo "$v" | tr ' =' '__')
  env $v timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --config 2 --steps 5 --warmup 3 --no-cpu-baseline --train-steps 16 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); t=d['train_step']
print('$v', 'train ms', round(t['ms_per_step'],3), 'videos/s', round(t['value']))" | tee -a gpurun_out/nccl_variants_n$N.txt
done
