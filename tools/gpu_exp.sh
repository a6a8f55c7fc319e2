#!/bin/bash
# GPU experiment driver: parity tests on the product library, then kernel-only timing of every tuning variant.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/exp_pytest.txt
: > gpurun_out/exp_perf.txt
timeout 120 python tools/quick_perf.py 62000 40 2>&1 | tail -1 | sed -e "s/^/default: /" | tee -a gpurun_out/exp_perf.txt
for v in cable_b200/variants/*.so; do
  CABLE_B200_LIB=$v python tools/quick_perf.py 62000 40 2>&1 | tail -1 | sed -e "s|^|$v: |" | tee -a gpurun_out/exp_perf.txt
done
