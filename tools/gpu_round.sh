#!/bin/bash
# One GPU visit: parity suite on the product library, kernel-only timing of the product and of every tuning variant
# under cable_b200/variants, then the default bench line.  Every step is bounded.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader | tee gpurun_out/gpu.txt
timeout -s KILL 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.txt
: > gpurun_out/exp_perf.txt
for r in 1 2; do timeout -s KILL 120 python tools/quick_perf.py 62000 60 2>&1 | tail -1 | sed -e "s/^/product: /" | tee -a gpurun_out/exp_perf.txt; done
for v in cable_b200/variants/*.so; do
  [ -f "$v" ] || continue
  if CABLE_B200_LIB=$v timeout -s KILL 60 python tools/quick_perf.py 7750 10 > /tmp/qp_small.txt 2>&1; then
    for r in 1 2; do CABLE_B200_LIB=$v timeout -s KILL 120 python tools/quick_perf.py 62000 60 2>&1 | tail -1 | sed -e "s|^|$v: |" | tee -a gpurun_out/exp_perf.txt; done
  else
    echo "$v: small run failed or timed out" | tee -a gpurun_out/exp_perf.txt
  fi
done
timeout -s KILL 600 python bench.py 2> gpurun_out/bench_err.txt | tee gpurun_out/bench_n1.json
tail -3 gpurun_out/bench_err.txt
