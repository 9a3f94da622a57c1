"""CPU: the C-ABI shared library builds, loads, exports every symbol of include/pagk.h, agrees with the
ctypes mirror on struct layout, and fails loudly without a device (no compute calls without a GPU)."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

from pixel_aware_gyro_aided_klt_feature_tracker_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_are_exported(cuda_lib):
    hdr = open(os.path.join(ROOT, "include", "pagk.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(pagk_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(capi.SYMBOLS), declared ^ set(capi.SYMBOLS)
    for name in declared:
        assert hasattr(cuda_lib, name), name
    assert cuda_lib.pagk_version() == 100


def test_struct_layout_matches_ctypes(tmp_path):
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "pagk.h"\nint main(void){\n'
                   'printf("%zu %zu %zu %zu %zu\\n", sizeof(pagk_config), sizeof(pagk_params), sizeof(pagk_pair_in), sizeof(pagk_pair_out), sizeof(pagk_patch_match_in));\n'
                   'printf("%zu %zu %zu %zu\\n", offsetof(pagk_pair_in, K), offsetof(pagk_pair_in, Rcl_override), offsetof(pagk_pair_out, Rcl), offsetof(pagk_pair_out, n_iterations));\n'
                   'printf("%zu %zu %zu %zu\\n", sizeof(pagk_geometry_in), sizeof(pagk_geometry_out), offsetof(pagk_geometry_in, estimate), offsetof(pagk_geometry_out, H21));\n'
                   'return 0;}\n')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = subprocess.check_output([str(exe)]).decode().split()
    sizes = [C.sizeof(x) for x in (capi.PagkConfig, capi.PagkParams, capi.PagkPairIn, capi.PagkPairOut, capi.PagkPatchMatchIn)]
    offs = [capi.PagkPairIn.K.offset, capi.PagkPairIn.Rcl_override.offset, capi.PagkPairOut.Rcl.offset,
            capi.PagkPairOut.n_iterations.offset]
    geo = [C.sizeof(capi.PagkGeometryIn), C.sizeof(capi.PagkGeometryOut), capi.PagkGeometryIn.estimate.offset, capi.PagkGeometryOut.H21.offset]
    assert [int(x) for x in out] == sizes + offs + geo


def test_default_params_are_the_reference_constants(cuda_lib):
    p = capi.PagkParams()
    cuda_lib.pagk_default_params(C.byref(p))
    assert (p.e_type, p.predict_method, p.half_patch, p.iterations, p.pyramids, p.inverse, p.calc_ncc) == (4, 1, 5, 10, 3, 0, 0)
    assert (p.lambda_, p.alpha, p.max_distance) == (1.0, 0.5, 25)


def test_no_device_means_error_not_fallback(cuda_lib):
    if cuda_lib.pagk_device_count() > 0:
        pytest.skip("a CUDA device is visible")
    cfg = capi.PagkConfig(0, 64, 64, 8, 1, 8, 2, 5)
    h = C.c_void_p()
    assert cuda_lib.pagk_create(C.byref(cfg), C.byref(h)) == capi.PAGK_ERR_NO_DEVICE
    assert not h.value and b"no CPU fallback" in cuda_lib.pagk_last_error()
    from pixel_aware_gyro_aided_klt_feature_tracker_b200 import tracker
    with pytest.raises(tracker.PagkError):
        tracker.Context()


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        capi.load(str(tmp_path / "libpagk_cuda.so"))


def test_product_package_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "pixel_aware_gyro_aided_klt_feature_tracker_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(d, f)).read()
                code = "\n".join(l for l in txt.splitlines() if not l.strip().startswith(("//", "#", "*", '"""')))
                assert "import oracle" not in code and "from oracle" not in code and "libpagk_oracle" not in code, f
