#!/bin/bash
# GPU experiment driver: parity tests on the product library, bit-identity digests and kernel-only timing of every tuning variant.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/exp_pytest.txt
: > gpurun_out/exp_perf.txt
run() { # label, env...
  local label=$1; shift
  env "$@" timeout 200 python tools/state_hash.py 62000 12 2>&1 | tail -1 | sed -e "s/^/$label: /" | tee -a gpurun_out/exp_perf.txt
  for r in 1 2; do env "$@" timeout 120 python tools/quick_perf.py 62000 60 2>&1 | tail -1 | sed -e "s/^/$label: /" | tee -a gpurun_out/exp_perf.txt; done
}
run default X=1
for v in cable_b200/variants/*.so; do
  [ -f "$v" ] || continue
  run "$(basename $v)" CABLE_B200_LIB=$v
done
