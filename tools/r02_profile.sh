#!/bin/bash
# Round-2 evidence for profiles/: compute-sanitizer (memcheck + racecheck) over the hand-rolled-protocol kernels, the ncu
# launch list of the bench step and full-set captures of its kernels.
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  for tgt in netvlad gemm lstm; do
    echo "=== compute-sanitizer --tool $tool: $tgt"
    timeout 600 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_targets.py $tgt 2>&1 | grep -v "^$" | tail -14
  done
done 2>&1 | tee gpurun_out/r02_sanitizer.log
echo "=== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 160 --csv --log-file gpurun_out/r02_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --train-steps 0 > gpurun_out/r02_launches_run.log 2>&1
tail -2 gpurun_out/r02_launches_run.log
echo "=== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"netvlad|gemm_tcgen05|frames_unpack|l2norm_rows" -s 6 -c 4 -f -o gpurun_out/r02_prof \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --train-steps 0 > gpurun_out/r02_prof_run.log 2>&1
tail -2 gpurun_out/r02_prof_run.log
ls -la gpurun_out/ | grep r02_
