#!/bin/bash
# N-rank data-parallel checks on one box: gradient parity vs a single process, then the train-step scaling lines.
# usage: tools/r02_dp.sh N
cd "$(dirname "$0")/.." || exit 1
N=${1:-2}
mkdir -p gpurun_out
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 "$@"; }
run tools/dp_train_check.py 2>&1 | grep -v Warning | tail -6 | tee gpurun_out/dp_train_check_n$N.txt
timeout 300 python bench.py --config 2 --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | tee gpurun_out/dp_bench_c2_n1.json | python -c "
import json,sys; d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); print('N=1', d['value'], d['train_step'], 'e2e', d['e2e']['value'])"
run bench.py --gpus $N --config 2 --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/dp_bench_c2_n$N.err | tee gpurun_out/dp_bench_c2_n$N.json | python -c "
import json,sys; d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); print('N=$N', d['value'], d['train_step'], 'e2e', d['e2e']['value'])"
tail -3 gpurun_out/dp_bench_c2_n$N.err
