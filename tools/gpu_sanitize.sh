#!/bin/bash
# One bounded GPU visit under compute-sanitizer: memcheck, racecheck (kernel B shares a tile's column between six threads
# through shared memory; kernel A stages parameter tables there), initcheck and synccheck of tools/sanitize_case.py.
# Logs go to gpurun_out/sanitize_<tool>.log; the last lines of each hold the tool's error summary.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NL=${1:-300}
for tool in memcheck racecheck synccheck initcheck; do
  extra=""
  [ "$tool" = memcheck ] && extra="--leak-check full"
  timeout -s KILL ${SAN_TIMEOUT:-55} compute-sanitizer --tool $tool $extra --print-limit 20 python tools/sanitize_case.py $NL 2 \
    > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|LEAK SUMMARY' gpurun_out/sanitize_$tool.log | tr '\n' ' ')"
  grep -E "^drop-in|^driver" gpurun_out/sanitize_$tool.log
done
