#!/bin/bash
# stall reasons of one LK launch for a library variant: tools/ncu_stalls.sh <variant>
cd "$(dirname "$0")/.."
v=$1
if [ "$v" != prod ]; then export PAGK_LIB=$PWD/tools/libpagk_$v.so; fi
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active
for r in long_scoreboard short_scoreboard wait not_selected no_instruction dispatch_stall lg_throttle mio_throttle math_pipe_throttle membar branch_resolving sleeping barrier; do M=$M,smsp__average_warps_issue_stalled_${r}_per_issue_active.ratio; done
M=$M,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,sm__icc_requests_lookup_miss.sum,sm__icc_requests.sum,smsp__inst_executed_pipe_lsu.sum
timeout 200 ncu --metrics $M --clock-control none -k regex:lk_lanes -s 6 -c 1 --csv --log-file gpurun_out/stalls_$v.csv python bench.py --steps 4 --warmup 3 --no-cpu > /dev/null 2> gpurun_out/stalls_$v.err
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/stalls_$v.csv")) if len(r)>5]
h=rows[0]; n=h.index("Metric Name"); val=h.index("Metric Value")
print("== $v")
for r in rows[1:]: print("  %-85s %s"%(r[n],r[val]))
PY
