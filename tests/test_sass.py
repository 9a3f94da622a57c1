"""CPU: the machine code of the built library is what the bit-parity argument of csrc/pagk_lk_lanes.cu assumes.

The pass of the alignment kernel is written with packed FP32 PTX (add/mul/fma.f32x2).  ptxas contracts a packed
multiply and a packed add into FFMA2 even under --fmad=false, so the source writes every "product + product" as
fma2(product, ONE, product) and the ONLY fused operations of the hot loop are the ones written out.  This test
disassembles the library (cuobjdump, no GPU needed) and counts them, so that a compiler change that fuses anything
else fails here before it fails the parity tests on the GPU.  It also checks that the data-movement instructions
the design names are in the kernels that are supposed to have them."""
import collections
import os
import re
import shutil
import subprocess

import pytest

from pixel_aware_gyro_aided_klt_feature_tracker_b200 import _build

CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
pytestmark = pytest.mark.skipif(not os.path.exists(CUOBJDUMP), reason="cuobjdump not installed")


@pytest.fixture(scope="module")
def kernels():
    """{mangled kernel name: [SASS instruction text, ...]} of csrc/libpagk_cuda.so"""
    txt = subprocess.run([CUOBJDUMP, "-sass", _build.build()], capture_output=True, text=True, check=True).stdout
    out = {}
    for f in re.split(r"\n\s*Function : ", txt)[1:]:
        name = f.split("\n", 1)[0].strip()
        ins = []
        for line in f.split("\n"):
            m = re.match(r"\s*/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
            if m:
                ins.append((int(m.group(1), 16), re.sub(r"^@!?U?P\d\s+", "", m.group(2))))
        out[name] = ins
    return out


def hot_loop(ins):
    """the shortest backward-branch body that holds the double-precision accumulation and the byte taps"""
    addr = {a: i for i, (a, _) in enumerate(ins)}
    best = None
    for i, (a, t) in enumerate(ins):
        m = re.search(r"BRA.*0x([0-9a-f]+)", t)
        if not m:
            continue
        tgt = int(m.group(1), 16)
        if tgt < a and tgt in addr:
            body = [x for _, x in ins[addr[tgt]:i + 1]]
            if sum("DFMA" in x for x in body) >= 16 and sum("LDS.U8" in x for x in body) >= 24:
                if best is None or len(body) < len(best):
                    best = body
    return best


def opcodes(body):
    return collections.Counter(x.split()[0].split(".")[0] for x in body)


@pytest.mark.parametrize("affine", [True, False])
def test_fused_operations_of_the_pass_are_the_written_ones(kernels, affine):
    names = [n for n in kernels if "pagk_lk_lanes_kernelILi5ELb%d" % int(affine) in n]
    assert len(names) >= 1, list(kernels)[:5]
    for n in names:
        body = hot_loop(kernels[n])
        assert body is not None, n
        c = opcodes(body)
        pairs = c["DFMA"] // 16        # a pixel pair has 2 x 8 DFMA
        # three conversions per pixel (a register-starved instantiation may rematerialise the loop-invariant c or gain)
        assert pairs in (2, 4) and c["DADD"] == 6 * pairs and 6 * pairs <= c["F2F"] <= 6 * pairs + 2, (n, c)
        # per pair: 13 interpolations + the residual's two + the two sample coordinates (affine: the warp offsets), each ONE
        # fma2 in the source -- fma2(x, 1, y), fma(v, 2^-51, db), fma(offset, 2^100, base) --; every other packed
        # operation must still be a separate FMUL2 / FADD2
        assert c["FFMA2"] == 17 * pairs, (n, c)
        assert c["FMUL2"] == (32 if affine else 28) * pairs, (n, c)
        # (a contraction would move one FMUL2 and one FADD2 into an FFMA2: the two counts above exclude it; an instantiation at
        # its register cap may recompute a loop-invariant packed constant inside the loop, which adds FADD2 only)
        assert (20 if affine else 18) * pairs <= c["FADD2"] <= (20 if affine else 18) * pairs + 2, (n, c)
        assert c["FFMA"] == 0 and c["FMUL"] == 0, (n, c)    # no scalar FP32 product left in the loop, fused or not
        assert c["LDS"] == 24 * pairs, (n, c)               # twelve byte taps per pixel, nothing spilled


def test_copy_engines_are_where_the_design_says(kernels):
    tmpl = [n for n in kernels if "pagk_lk_template_kernel" in n]
    lanes = [n for n in kernels if "pagk_lk_lanes_kernel" in n]
    assert tmpl and lanes
    for n in tmpl:   # K3a: one TMA tile load per work item, completion on an mbarrier
        ops = " ".join(x for _, x in kernels[n])
        assert "UTMALDG" in ops and "SYNCS" in ops, n
    for n in lanes:  # K3b: windows by asynchronous copies straight into shared memory, no register staging
        ops = " ".join(x for _, x in kernels[n])
        assert "LDGSTS" in ops and "LDGDEPBAR" in ops, n
