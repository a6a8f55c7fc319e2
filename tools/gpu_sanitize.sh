#!/bin/bash
# One bounded GPU visit under compute-sanitizer: memcheck, racecheck (kernel B shares a tile's column between six threads
# through shared memory; kernel A stages parameter tables there), initcheck and synccheck of tests/checks/sanitize_case.py.
# SAN_TOOLS / SAN_STEPS / SAN_TAG / SAN_TIMEOUT select tools, steps, a log suffix and the per-tool bound (a second visit at
# > 148 x 768 tiles covers the full-rounds + remainder launch chains).  Logs go to gpurun_out/sanitize_<tool>.log; the last lines of each hold the tool's error summary.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NL=${1:-300}
for tool in ${SAN_TOOLS:-memcheck racecheck synccheck initcheck}; do
  extra=""
  [ "$tool" = memcheck ] && extra="--leak-check full"
  timeout -s KILL ${SAN_TIMEOUT:-55} compute-sanitizer --tool $tool $extra --print-limit 20 python tests/checks/sanitize_case.py $NL ${SAN_STEPS:-2} \
    > gpurun_out/sanitize_${tool}${SAN_TAG}.log 2>&1
  echo "== $tool rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|LEAK SUMMARY' gpurun_out/sanitize_${tool}${SAN_TAG}.log | tr '\n' ' ')"
  grep -E "^drop-in|^driver|Error|error" gpurun_out/sanitize_${tool}${SAN_TAG}.log
done
