M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__occupancy_limit_registers,launch__occupancy_limit_shared_mem,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed
for r in long_scoreboard short_scoreboard wait not_selected no_instruction dispatch_stall lg_throttle mio_throttle math_pipe_throttle membar branch_resolving sleeping barrier; do M=$M,smsp__average_warps_issue_stalled_${r}_per_issue_active.ratio; done
timeout 300 ncu --metrics $M --clock-control none -k regex:"template|pyramid" -s 6 -c 2 --csv --log-file gpurun_out/k3a.csv python bench.py --steps 4 --warmup 3 --no-cpu > /dev/null 2> gpurun_out/k3a.err
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/k3a.csv")) if len(r)>5]
h=rows[0]; n=h.index("Metric Name"); val=h.index("Metric Value"); k=h.index("Kernel Name")
for r in rows[1:]: print("%-28s %-80s %s"%(r[k][:28], r[n],r[val]))
PY
