import numpy as np


def assert_bitwise(a, b, what=""):
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    assert a.dtype == b.dtype == np.float32, f"{what}: dtype {a.dtype} vs {b.dtype}"
    if not np.array_equal(a, b, equal_nan=True):
        bad = ~((a == b) | (np.isnan(a) & np.isnan(b)))
        idx = np.argwhere(bad)[0]
        raise AssertionError(f"{what}: {bad.sum()} of {a.size} elements differ, first at {tuple(idx)}: "
                             f"{a[tuple(idx)]!r} vs {b[tuple(idx)]!r}")


def rel_err(a, b):
    a = a.astype(np.float64)
    b = b.astype(np.float64)
    assert np.array_equal(np.isnan(a), np.isnan(b)), "NaN patterns differ"
    d = np.abs(a - b)
    return float(np.nanmax(d) / (np.nanmax(np.abs(b)) + 1e-30)) if d.size else 0.0


def mean_epe(u0, v0, u1, v1):
    """Mean end-point error in pixels between two flow fields."""
    return float(np.mean(np.sqrt((u0.astype(np.float64) - u1) ** 2 + (v0.astype(np.float64) - v1) ** 2)))
