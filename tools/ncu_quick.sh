#!/bin/bash
# quick instruction-supply profile of the two step kernels: bash tools/ncu_quick.sh <tag> [lib.so]
cd "$(dirname "$0")/.."
tag=$1; lib=${2:-}
M=gpu__time_duration.sum,smsp__inst_executed.sum,sm__icc_requests.sum,sm__icc_request_hit_rate.pct,gcc__cache_requests_type_instruction.sum,gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,sm__warps_active.avg.pct_of_peak_sustained_active,launch__grid_size,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__average_warp_latency_per_inst_issued.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed_op_branch.sum,l1tex__t_sector_hit_rate.pct
CABLE_B200_LIB=$lib ncu --metrics $M -k regex:cbm_kernel -s 20 -c 2 --csv --log-file gpurun_out/ncuq_$tag.csv python tools/quick_perf.py 62000 12 > gpurun_out/ncuq_$tag.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/ncuq_$tag.csv")) if len(r)>10]
h=rows[0]; 
for r in rows[1:]:
    print("$tag", r[h.index("Kernel Name")][:28], r[h.index("Metric Name")], r[h.index("Metric Value")])
PY
