"""In-tree nvcc build of csrc/libpagk_cuda.so for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
SOURCES = ["pagk_kernels.cu", "pagk_lk_lanes.cu", "pagk_api.cu"]
HEADERS = ["pagk_device.cuh", "pagk_kernels.h", "pagk_host_math.h", "pagk_octree.h", "pagk_ransac.h", os.path.join("..", "..", "include", "pagk.h")]
OUT = os.path.join(CSRC, "libpagk_cuda.so")
# the same library with the index traps of the LK lanes kernel compiled in (-DPAGK_LANES_CHECK): test infrastructure,
# selected with PAGK_LIB=<this path>; compute-sanitizer is not available on the GPU pool
OUT_CHECK = os.path.join(CSRC, "libpagk_cuda_check.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--fmad=false",                       # bit-parity with the reference's SSE2 (no FMA) build
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math,-O2",
    "-shared", "-cudart", "static",
]


def nvcc_path() -> str:
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found: the pagk CUDA library cannot be built (there is no CPU fallback)")
    return p


def needs_build(out: str = OUT) -> bool:
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, check: bool = False) -> str:
    out = OUT_CHECK if check else OUT
    if not force and not needs_build(out):
        return out
    extra = os.environ.get("PAGK_NVCC_EXTRA", "").split() + (["-DPAGK_LANES_CHECK"] if check else [])
    cmd = [nvcc_path()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + \
          ["-o", out] + [os.path.join(CSRC, s) for s in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stderr[-4000:])
    return out


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, check="--check" in sys.argv))
