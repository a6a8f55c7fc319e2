#!/bin/bash
# small-range geometry of kernel A (128-thread blocks, CBL_SMALL_MINB per SM) at strong-scaling shard sizes
cd "$(dirname "$0")/.."
out=gpurun_out/${1:-r02_small_minb_probe}.txt; : > $out
for nland in 7750 11400 15500 19000; do
  for v in default cable_b200/variants/s2.so cable_b200/variants/s4.so cable_b200/variants/s5.so cable_b200/variants/s6.so; do
    if [ $v = default ]; then unset CABLE_B200_LIB; else export CABLE_B200_LIB=$v; fi
    echo -n "$v (small geometry forced): " | tee -a $out
    CABLE_B200_BIG_MIN=400000 timeout -s KILL 120 python tools/quick_perf.py $nland 40 2>&1 | tail -1 | tee -a $out
  done
  unset CABLE_B200_LIB
  echo -n "default (big geometry forced): " | tee -a $out
  CABLE_B200_BIG_MIN=1 timeout -s KILL 120 python tools/quick_perf.py $nland 40 2>&1 | tail -1 | tee -a $out
done
