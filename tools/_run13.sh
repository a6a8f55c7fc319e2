cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:cbm_kernel -s 36 -c 3 -f -o gpurun_out/r02_step_full python tools/quick_perf.py 62000 12 > gpurun_out/r02_step_full.log 2>&1
ls -la gpurun_out/*.ncu-rep
