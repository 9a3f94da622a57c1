"""Shared comparison helpers for the parity tests (CUDA path vs CPU oracle)."""
import numpy as np

FLOAT_FIELDS = ["pt_predict_un", "pt_predict", "pt_gyro_predict_un", "pt_gyro_predict", "flows_predict_un", "affine",
                "corner_flows", "pt_corners_un", "pt_corners", "pm_pt_un", "pm_pt", "pixel_error", "distance", "ncc"]
INT_FIELDS = ["status", "pm_status", "iters"]


def bits_equal(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """element-wise equality of the bit patterns (NaN == NaN, +0 != -0)"""
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    u = {4: np.uint32, 8: np.uint64, 1: np.uint8}[a.dtype.itemsize]
    return a.view(u) == b.view(u)


def compare(gpu, cpu):
    """dict of per-field mismatch statistics between two PairOutputs"""
    rep = {}
    for f in FLOAT_FIELDS:
        g, c = getattr(gpu, f), getattr(cpu, f)
        eq = bits_equal(g, c)
        both_nan = np.isnan(g) & np.isnan(c)
        with np.errstate(invalid="ignore"):
            d = np.where(both_nan, 0.0, np.abs(g.astype(np.float64) - c.astype(np.float64)))
        rep[f] = dict(bit_mismatch=int((~eq).sum()), max_abs=float(np.nanmax(d)) if d.size else 0.0,
                      n=int(g.size))
    for f in INT_FIELDS:
        g, c = getattr(gpu, f), getattr(cpu, f)
        rep[f] = dict(bit_mismatch=int((g != c).sum()), n=int(g.size))
    rep["n_predict"] = dict(gpu=gpu.n_predict, cpu=cpu.n_predict)
    rep["n_iterations"] = dict(gpu=gpu.n_iterations, cpu=cpu.n_iterations)
    rep["Rcl_bits"] = int((~bits_equal(gpu.Rcl, cpu.Rcl)).sum())
    rep["KRKinv_bits"] = int((~bits_equal(gpu.KRKinv, cpu.KRKinv)).sum())
    return rep


def assert_north_star(gpu, cpu, pos_tol=0.01, status_frac=0.999):
    """BASELINE.json north_star: tracked positions within 0.01 px, identical status on >= 99.9 %"""
    n = max(1, cpu.status.size)
    same = (gpu.status == cpu.status)
    assert same.mean() >= status_frac or (~same).sum() == 0, f"status differs on {(~same).sum()}/{n}"
    ok = (cpu.status == 1) & same
    if ok.any():
        for f in ("pt_predict_un", "pt_predict"):
            d = np.hypot(*(getattr(gpu, f)[ok] - getattr(cpu, f)[ok]).T)
            assert d.max() <= pos_tol, f"{f}: max position error {d.max()} px > {pos_tol}"


def assert_bit_exact(gpu, cpu, fields=None):
    rep = compare(gpu, cpu)
    bad = {k: v for k, v in rep.items() if isinstance(v, dict) and v.get("bit_mismatch", 0) and (fields is None or k in fields)}
    assert not bad, f"bit mismatches: {bad}"
    if fields is None:
        assert rep["Rcl_bits"] == 0 and rep["KRKinv_bits"] == 0, rep
        assert gpu.n_predict == cpu.n_predict and gpu.n_iterations == cpu.n_iterations, rep
