#!/bin/bash
# limiter metrics of one kernel: tools/ncu_kernel.sh <kernel regex> <python script and args...>
cd "$(dirname "$0")/.."
k=$1; shift
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed,lts__t_sectors_srcunit_tex_op_read.sum,l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed,l1tex__f_wavefronts.sum.pct_of_peak_sustained_elapsed
for r in long_scoreboard short_scoreboard wait not_selected no_instruction dispatch_stall lg_throttle mio_throttle tex_throttle math_pipe_throttle barrier; do M=$M,smsp__average_warps_issue_stalled_${r}_per_issue_active.ratio; done
timeout 300 ncu --metrics $M --clock-control none -k regex:"$k" -s 2 -c 1 --csv --log-file gpurun_out/ncu_kernel.csv python "$@" > /dev/null 2> gpurun_out/ncu_kernel.err
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/ncu_kernel.csv")) if len(r)>5]
h=rows[0]; n=h.index("Metric Name"); val=h.index("Metric Value")
for r in rows[1:]: print("  %-85s %s"%(r[n],r[val]))
PY
