#!/bin/bash
# kernel-only timing of every tuning variant under cable_b200/variants (no tests).  Each run is bounded with SIGKILL:
# a variant that deadlocks must not hold the GPU box until gpurun's own limit.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; : > gpurun_out/exp_perf.txt
for v in cable_b200/variants/*.so; do
  if CABLE_B200_LIB=$v timeout -s KILL 60 python tools/quick_perf.py 7750 10 > /tmp/qp_small.txt 2>&1; then
    CABLE_B200_LIB=$v timeout -s KILL 120 python tools/quick_perf.py 62000 40 2>&1 | tail -1 | sed -e "s|^|$v: |" | tee -a gpurun_out/exp_perf.txt
  else
    echo "$v: small run failed or timed out" | tee -a gpurun_out/exp_perf.txt
  fi
done
