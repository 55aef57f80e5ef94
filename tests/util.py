import numpy as np


def parity(got, ref, scale):
    """SURVEY.md H7 metric: max |got-ref| / max(|ref|, scale) with identical NaN patterns."""
    got = np.asarray(got, dtype=float)
    ref = np.asarray(ref, dtype=float)
    np.testing.assert_array_equal(np.isnan(got), np.isnan(ref), err_msg="NaN pattern differs")
    ok = ~np.isnan(ref)
    if not ok.any():
        return 0.0
    return float(np.max(np.abs(got[ok] - ref[ok]) / np.maximum(np.abs(ref[ok]), scale)))


def step_scale(row):
    s = np.nanmax(np.abs(row))
    return float(s) if np.isfinite(s) and s > 0 else 1.0


TOL = 1e-9          # north_star: 1e-9 relative in fp64
