#!/usr/bin/env python
"""Instruction mix of the hot loop of a kernel in a cubin/.so: tools/sass_loop.py <lib> <kernel substring>
Finds the backward branch with the most DFMA between target and branch and prints the per-opcode counts."""
import re, subprocess, sys, collections
lib, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)
for f in funcs[1:]:
    name = f.split("\n", 1)[0]
    if pat not in name:
        continue
    ins = []
    for line in f.split("\n"):
        m = re.match(r"\s*/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2)))
    addr2i = {a: i for i, (a, _) in enumerate(ins)}
    best = None
    for i, (a, t) in enumerate(ins):
        m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?`?\(?\.?L?_?x?_?\d*\)?\s*0x([0-9a-f]+)", t)
        m2 = re.search(r"BRA.*0x([0-9a-f]+)", t)
        if m2:
            tgt = int(m2.group(1), 16)
            if tgt < a and tgt in addr2i:
                body = ins[addr2i[tgt]:i + 1]
                nd = sum(1 for _, x in body if "DFMA" in x or "DADD" in x)
                nl = sum(1 for _, x in body if "LDS.U8" in x or "LDS.U16" in x)
                if nd >= 10 and nl >= 12 and (best is None or len(body) < len(best[1])):
                    best = (nd, body)
    print(name, "total instr", len(ins), "bytes", len(ins) * 16)
    if best:
        body = best[1]
        c = collections.Counter()
        for _, x in body:
            x = re.sub(r"^@!?U?P\d\s+", "", x)
            c[x.split()[0].split(".")[0]] += 1
        print(" hot loop:", len(body), "instr,", best[0], "FP64 acc ops")
        print(" ", dict(c.most_common()))
