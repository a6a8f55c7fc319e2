#!/bin/bash
# kernel-A block / register-cap variants x chain streams (chunk = one round of kernel A blocks per SM, the default)
cd "$(dirname "$0")/.."
out=gpurun_out/${1:-r02_pipe_probe2}.txt; : > $out
for v in default cable_b200/variants/*.so; do
  for s in 0 3 4 6; do
    echo -n "$v streams=$s: " | tee -a $out
    if [ $v = default ]; then unset CABLE_B200_LIB; else export CABLE_B200_LIB=$v; fi
    CABLE_B200_PIPE_STREAMS=$s timeout -s KILL 120 python tools/quick_perf.py 62000 40 2>&1 | tail -1 | tee -a $out
  done
done
