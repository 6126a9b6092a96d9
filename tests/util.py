"""Shared helpers for the parity tests."""
import numpy as np


def vec(a):
    return np.stack([a["x"], a["y"], a["z"]], axis=1)


def rel_err(a, ref, mask=None):
    """Per-body |a - ref| / |ref| (vector norms)."""
    a, ref = vec(a), vec(ref)
    if mask is not None:
        a, ref = a[mask], ref[mask]
    num = np.linalg.norm(a - ref, axis=1)
    den = np.linalg.norm(ref, axis=1)
    return num / np.where(den > 0, den, 1.0)


# Stated tolerance of BASELINE.json's north_star for fp32 device forces vs the fp64 reference.
MEDIAN_TOL = 1e-5
P99_TOL = 1e-3


def assert_acc_parity(a, ref, mask=None, median=MEDIAN_TOL, p99=P99_TOL):
    r = rel_err(a, ref, mask)
    assert np.isfinite(r).all()
    assert np.median(r) <= median, f"median rel err {np.median(r):.3e}"
    assert np.percentile(r, 99) <= p99, f"p99 rel err {np.percentile(r, 99):.3e}"
    return r
