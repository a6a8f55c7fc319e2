#!/bin/bash
# One GPU visit for the committed profiles: ncu launch list of the bench command, one --set full capture of one
# resident step (kernel A fast build, ordinary kernel A over the handed-back blocks, kernel B; two chains each).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 2 --warmup 1 > gpurun_out/launches_bench.log 2>&1
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:cbm_kernel -s 36 -c 6 -f -o gpurun_out/step_full \
  python tools/quick_perf.py 62000 12 > gpurun_out/step_full.log 2>&1
ls -la gpurun_out/ | tail -8
