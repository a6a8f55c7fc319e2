#!/bin/bash
# mid-size shards (strong scaling at N = 2, 4, 8): kernel A geometry threshold
cd "$(dirname "$0")/.."
out=gpurun_out/${1:-r02_mid_probe}.txt; : > $out
for nland in 31000 15500 7750 20000 11400; do
  for bm in 94720 56833 28000; do
    echo -n "big_min=$bm: " | tee -a $out
    CABLE_B200_BIG_MIN=$bm timeout -s KILL 120 python tools/quick_perf.py $nland 40 2>&1 | tail -1 | tee -a $out
  done
done
