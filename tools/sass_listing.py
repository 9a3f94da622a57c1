"""SASS evidence for profiles/: per kernel of the library, the counts of the data-movement and packed-arithmetic
instructions the design names, the lines that hold them in K3a, and the hot loop of the production alignment kernel.
usage: sass_listing.py <lib.so> <out.txt>"""
import collections, re, subprocess, sys
import os
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
lib, out = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
MARK = ["UTMALDG", "UTMASTG", "UBLKCP", "LDGSTS", "LDGDEPBAR", "SYNCS", "FFMA2", "FMUL2", "FADD2", "DFMA", "F2F", "UTCHMMA", "LDTM"]
with open(out, "w") as f:
    f.write(f"# cuobjdump -sass {lib.split('/')[-1]} (sm_100a), summarised by tools/sass_listing.py\n\n")
    f.write("## instruction counts per kernel (static)\n\n%-64s %6s " % ("kernel", "total") + " ".join("%9s" % m for m in MARK) + "\n")
    funcs = re.split(r"\n\s*Function : ", txt)[1:]
    parsed = {}
    for fn in funcs:
        name = fn.split("\n", 1)[0].strip()
        ins = [(int(m.group(1), 16), m.group(2)) for m in (re.match(r"\s*/\*([0-9a-f]{4,5})\*/\s+(.*?);", l) for l in fn.split("\n")) if m]
        parsed[name] = ins
        dem = subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip().split("(")[0].replace("void ", "")
        c = collections.Counter()
        for _, t in ins:
            op = re.sub(r"^@!?U?P\d\s+", "", t).split()[0]
            for m in MARK:
                if op.startswith(m):
                    c[m] += 1
        f.write("%-64s %6d " % (dem[:64], len(ins)) + " ".join("%9d" % c[m] for m in MARK) + "\n")
    for name, ins in parsed.items():
        if "pagk_lk_template_kernelILi5" in name:
            f.write("\n## K3a pagk_lk_template_kernel<5>: the tile loads and their barrier\n\n")
            for a, t in ins:
                if any(k in t for k in ("UTMALDG", "SYNCS", "UTMACCTL", "FENCE.VIEW.ASYNC")):
                    f.write("  /*%04x*/  %s\n" % (a, t))
    for name, ins in parsed.items():
        if "pagk_lk_lanes_kernelILi5ELb1ELi8" in name:
            f.write("\n## K3b pagk_lk_lanes_kernel<5, true, 8>: the window copies\n\n")
            for a, t in ins:
                if any(k in t for k in ("LDGSTS", "LDGDEPBAR", "DEPBAR")):
                    f.write("  /*%04x*/  %s\n" % (a, t))
            import test_sass
            body = test_sass.hot_loop([(a, re.sub(r"^@!?U?P\d\s+", "", t)) for a, t in ins])
            addr = {a: i for i, (a, _) in enumerate(ins)}
            # locate the body again with addresses
            for i, (a, t) in enumerate(ins):
                m = re.search(r"BRA.*0x([0-9a-f]+)", t)
                if m and int(m.group(1), 16) < a and int(m.group(1), 16) in addr and i + 1 - addr[int(m.group(1), 16)] == len(body):
                    lo = addr[int(m.group(1), 16)]
                    f.write("\n## K3b pagk_lk_lanes_kernel<5, true, 8>: the pass loop (%d instructions for 8 pixels)\n\n" % len(body))
                    c = collections.Counter(re.sub(r"^@!?U?P\d\s+", "", t2).split()[0].split(".")[0] for _, t2 in ins[lo:i + 1])
                    f.write("  " + ", ".join("%s %d" % kv for kv in c.most_common()) + "\n\n")
                    for a2, t2 in ins[lo:i + 1]:
                        f.write("  /*%04x*/  %s\n" % (a2, t2))
                    break
print(open(out).read()[:3000])
