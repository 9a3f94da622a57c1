"""Per-CUDA-source-line executed-instruction and stall-sample shares of the profiled kernel, from an ncu capture
taken with --import-source on (kernel built with -lineinfo).   usage: ncu_lines.py report.ncu-rep [min_pct]"""
import csv, subprocess, sys
rep = sys.argv[1]; min_pct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass,cuda", "--csv"], capture_output=True, text=True).stdout
data, fname, hdr = [], "", None
for r in csv.reader(raw.splitlines()):
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No":
        hdr = r; ie = hdr.index("Instructions Executed"); isamp = hdr.index("# Samples"); continue
    if hdr is None or len(r) < len(hdr) or r[0] == "":
        continue
    try:
        data.append((fname, int(r[0]), r[1], int(r[ie].replace(",", "") or 0), int(r[isamp].replace(",", "") or 0)))
    except ValueError:
        pass
tot = sum(d[3] for d in data) or 1; tots = sum(d[4] for d in data) or 1
print(f"total warp-instructions {tot}, samples {tots}")
for f, ln, src, n, s in data:
    if 100.0 * n / tot >= min_pct or 100.0 * s / tots >= min_pct:
        print(f"{f[:18]:18s}{ln:5d} {100.0 * n / tot:6.2f}% inst {100.0 * s / tots:6.2f}% smp | {src.strip()[:110]}")
