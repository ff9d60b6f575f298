#!/bin/bash
# persistent-LSTM bring-up on the GPU box: each shape in its own process (a trapping kernel poisons only that process).
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
for s in "3 5 64 256 1" "70 9 128 256 2" "5 7 64 512 1" "64 12 1152 1024 2" "16 300 1152 1024 2"; do
  echo "=== $s"
  timeout 150 python tools/lstm_rec_check.py $s 2>&1 | tail -8
done 2>&1 | tee gpurun_out/lstm_rec_check.txt
echo "=== timing"
timeout 200 python tools/lstm_rec_check.py 64 300 1152 1024 2 time 2>&1 | tail -8 | tee -a gpurun_out/lstm_rec_check.txt
timeout 200 python tools/lstm_rec_check.py 128 300 1152 1024 2 time 2>&1 | tail -8 | tee -a gpurun_out/lstm_rec_check.txt
