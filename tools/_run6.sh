bash tools/gpu_small_variants.sh r02_small_variants_3
for v in cable_b200/variants/s128_4_inl.so; do CABLE_B200_LIB=$v python tools/quick_perf.py 62000 60 | tail -1 | tee -a gpurun_out/r02_small_variants_3.txt; done
python tools/quick_perf.py 62000 60 | tail -1 | tee -a gpurun_out/r02_small_variants_3.txt
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches_small.csv python tools/quick_perf.py 7750 6 > /dev/null 2>&1
