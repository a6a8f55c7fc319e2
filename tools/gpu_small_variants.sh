#!/bin/bash
# kernel-only timing of the small-shard geometry variants (cable_b200/variants/*.so) at the strong-scaling shard sizes:
# 7 750 / 15 500 / 31 000 land points = 38 750 / 77 500 / 155 000 tiles (config 3 at N = 8 / 4 / 2).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; out=gpurun_out/${1:-small_variants}.txt; : > $out
for n in 7750 15500 31000; do
  timeout -s KILL 120 python tools/quick_perf.py $n 60 2>&1 | tail -1 | sed -e "s|^|default: |" | tee -a $out
  for v in cable_b200/variants/*.so; do
    CABLE_B200_LIB=$v timeout -s KILL 120 python tools/quick_perf.py $n 60 2>&1 | tail -1 | sed -e "s|^|$v: |" | tee -a $out
  done
done
