cd "$(dirname "$0")/.."
out=gpurun_out/r02_variants_block2.txt; : > $out
for n in 62000 250000 100000; do
  for v in cable_b200/variants/b512.so cable_b200/variants/b512x2.so cable_b200/variants/b640.so cable_b200/variants/b896.so; do
    CABLE_B200_LIB=$v timeout -s KILL 120 python tools/quick_perf.py $n 40 2>&1 | tail -1 | sed -e "s|^|$v: |" | tee -a $out
  done
done
