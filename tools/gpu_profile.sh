#!/bin/bash
# One GPU visit for the committed profiles: ncu launch list of the bench command, one --set full capture of one
# resident step.  With the pipelined step (4 chunk chains at 310 000 tiles) a step is 12 launches: per chunk kernel A's
# CBL_FASTDIV build, the ordinary kernel A over the blocks that build handed back (normally none), kernel B.
# usage: tools/gpu_profile.sh [tag]      -> gpurun_out/<tag>_launches.csv, <tag>_step_full.ncu-rep
cd "$(dirname "$0")/.."
tag=${1:-r02}
mkdir -p gpurun_out
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/${tag}_launches_bench.log 2>&1
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:cbm_kernel -s 96 -c 12 -f -o gpurun_out/${tag}_step_full \
  python tools/quick_perf.py 62000 12 > gpurun_out/${tag}_step_full.log 2>&1
ls -la gpurun_out/ | tail -8
