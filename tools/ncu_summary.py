"""Summarise an ncu launch list (csv) and full captures (.ncu-rep) into a markdown file (run where ncu exists).
usage: ncu_summary.py <launches.csv> <out.md> <capture.ncu-rep> [<capture.ncu-rep> ...]"""
import collections, csv, subprocess, sys

launch_csv, out, reps = sys.argv[1], sys.argv[2], sys.argv[3:]
rows = [r for r in csv.reader(open(launch_csv)) if len(r) > 5]
hdr = rows[0]; ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    try:
        agg.setdefault(r[ki].split("(")[0].replace("void ", ""), []).append(float(r[vi].replace(",", "")))
    except ValueError:
        pass
tot = sum(sum(v) for v in agg.values())
keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]
with open(out, "w") as f:
    f.write("## launch list: device time per kernel (ncu, cold cache, serialised: compare shares, not absolutes)\n\n")
    f.write("| kernel | launches | avg us | share |\n|---|---:|---:|---:|\n")
    for k, val in agg.items():
        f.write(f"| `{k}` | {len(val)} | {sum(val) / len(val) / 1e3:.1f} | {100 * sum(val) / tot:.1f} % |\n")
    for rep in reps:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rr = list(csv.reader(raw.splitlines())); h, u, v = rr[0], rr[1], rr[2]
        name = v[h.index("Kernel Name")] if "Kernel Name" in h else rep
        f.write(f"\n## `{name.split('(')[0]}`: `ncu --set full`, one launch ({rep.split('/')[-1]})\n\n| metric | unit | value |\n|---|---|---:|\n")
        for k in keys:
            if k in h:
                f.write(f"| {k} | {u[h.index(k)]} | {v[h.index(k)]} |\n")
print(open(out).read())
