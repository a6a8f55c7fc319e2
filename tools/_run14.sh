cd "$(dirname "$0")/.."
out=gpurun_out/r02_variants_sync.txt; : > $out
for n in 62000 7750; do
  timeout -s KILL 120 python tools/quick_perf.py $n 60 2>&1 | tail -1 | sed -e "s|^|default: |" | tee -a $out
  for v in cable_b200/variants/*.so; do
    CABLE_B200_LIB=$v timeout -s KILL 120 python tools/quick_perf.py $n 60 2>&1 | tail -1 | sed -e "s|^|$v: |" | tee -a $out
  done
done
