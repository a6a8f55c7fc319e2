#!/bin/bash
# kernel-only timing of every tuning variant under cable_b200/variants (no tests)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; : > gpurun_out/exp_perf.txt
for v in cable_b200/variants/*.so; do
  CABLE_B200_LIB=$v python tools/quick_perf.py 62000 40 2>&1 | tail -1 | sed -e "s|^|$v: |" | tee -a gpurun_out/exp_perf.txt
done
