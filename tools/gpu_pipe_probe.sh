#!/bin/bash
# resident-step pipeline probe: chain streams x chunk size (CABLE_B200_PIPE_STREAMS=0 switches the pipeline off)
cd "$(dirname "$0")/.."
out=gpurun_out/${1:-r02_pipe_probe}.txt; : > $out
for nland in 62000 250000 7750; do
  for cfg in "0 0" "2 94720" "3 94720" "4 94720" "3 47360" "4 47360" "3 189440" "2 189440"; do
    set -- $cfg
    echo -n "streams=$1 chunk=$2: " | tee -a $out
    CABLE_B200_PIPE_STREAMS=$1 CABLE_B200_PIPE_CHUNK=$2 timeout -s KILL 120 python tools/quick_perf.py $nland 40 2>&1 | tail -1 | tee -a $out
  done
done
