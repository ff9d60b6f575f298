#!/bin/bash
# 8-rank (or N-rank) data-parallel evidence on one box: gradient parity vs a single process, all-reduce micro-timings, train-step scaling.
cd "$(dirname "$0")/.." || exit 1
N=${1:-8}
mkdir -p gpurun_out
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 "$@"; }
last_json() { python -c "
import json,sys
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1])
print('$1', 'value', round(d['value']), 'train', d.get('train_step') and (round(d['train_step']['value']), round(d['train_step']['ms_per_step'],3)), 'e2e', round(d['e2e']['value']))"; }
echo "=== dp_train_check at $N ranks"
run tools/dp_train_check.py 2>&1 | grep -v "Warning\|OMP_NUM\|\*\*\*\*" | tail -6 | tee gpurun_out/dp_train_check_n$N.txt
echo "=== all-reduce micro-timings"
run tools/allreduce_probe.py 2>&1 | grep -v "Warning\|OMP_NUM\|\*\*\*\*" | tail -12 | tee gpurun_out/allreduce_probe_n$N.txt
echo "=== bench config 2, N = 1"
timeout 300 python bench.py --config 2 --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | tee gpurun_out/dp_bench_c2_n1.json | last_json N=1
echo "=== bench config 2, N = $N"
run bench.py --gpus $N --config 2 --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/dp_bench_c2_n$N.err | tee gpurun_out/dp_bench_c2_n$N.json | last_json N=$N
# NCCL settings that change how the gradient pieces share the GPU with the backward (DP_VARIANTS="NCCL_ALGO=NVLS NCCL_MAX_CTAS=8")
for v in ${DP_VARIANTS}; do
  echo "=== bench config 2, N = $N, $v"
  env $v timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --config 2 --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/dp_bench_c2_n${N}_$v.err | tee gpurun_out/dp_bench_c2_n${N}_$v.json | last_json "N=$N,$v"
  env $v timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tools/allreduce_probe.py 2>&1 | grep "fp32\|bf16" | tee gpurun_out/allreduce_probe_n${N}_$v.txt
done
echo "=== bench config 2, N = $N, bf16 gradient wire format"
run bench.py --gpus $N --config 2 --steps 20 --warmup 5 --no-cpu-baseline --dp-gradient-dtype bfloat16 2>/dev/null | tee gpurun_out/dp_bench_c2_n${N}_bf16wire.json | last_json "N=$N,bf16-wire"
echo "=== train step per-launch timeline at $N ranks"
run tools/train_step_profile.py 2>&1 | tail -75 > gpurun_out/train_profile_n$N.txt
grep "world" gpurun_out/train_profile_n$N.txt
for c in ${DP_EXTRA_CONFIGS}; do
  echo "=== bench config $c, N = $N"
  run bench.py --gpus $N --config $c --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/dp_bench_c${c}_n$N.err | tee gpurun_out/dp_bench_c${c}_n$N.json | last_json N=$N
done
