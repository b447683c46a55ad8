"""Cubic-in-time interpolation of the grid field exactly as the reference does it, in a form a GPU can use.

The reference fits one cubic interpolating spline per grid point and component with ``scipy.interpolate.splrep``
(gizmo_interface.py:587-597; k=3, s=0 -> interpolating, not-a-knot) and evaluates all 3*(Ngrid+1) of them with
``splev`` on every step (gizmo_interface.py:607-620).  All those splines share ONE knot vector (the snapshot times),
so ``splev(t, tck_i) = sum_j B_j(t) * c_j[i]`` with the same four non-zero basis values B_j(t) for every grid point:
the evaluation collapses to a blend of four coefficient planes — what ``ocg_grid_interp_multi`` fuses into the gather.

``fit`` is one vectorised ``make_interp_spline`` over all grid points (bit-close to splrep, pinned by
tests/golden/time_spline_reference.npz); ``basis`` is the Cox-de Boor recurrence (no scipy at evaluation time).
"""
import numpy as np


def fit(times, values):
    """times [Nsnap] (strictly increasing, Nsnap >= 4), values [Nsnap, ...] -> (knots [Nsnap+4], coef [Nsnap, ...])."""
    from scipy.interpolate import make_interp_spline
    times = np.asarray(times, np.float64)
    if times.shape[0] < 4:
        raise ValueError("cubic time interpolation needs >= 4 snapshots (splrep: m > k must hold); got %d" % times.shape[0])
    spl = make_interp_spline(times, np.asarray(values, np.float64), k=3)
    return np.asarray(spl.t, np.float64), np.asarray(spl.c, np.float64)


def basis(knots, x, k=3):
    """Non-zero B-spline basis functions at x: returns (first, w[k+1]) with
    spline(x) = sum_{j=0..k} w[j] * coef[first + j].  x is clamped to the interpolation interval: outside the snapshot
    range the spline is HELD at its end value (splev's default ext=0 would extrapolate the end polynomials instead; the
    reference never leaves the snapshot range, gizmo_interface.py:607-620, so the two never differ on its path)."""
    t = np.asarray(knots, np.float64)
    n = len(t) - k - 1
    x = float(min(max(x, t[k]), t[n]))
    # knot span: largest i in [k, n-1] with t[i] <= x
    i = int(np.searchsorted(t, x, side="right")) - 1
    i = min(max(i, k), n - 1)
    w = np.zeros(k + 1)
    w[0] = 1.0
    left = np.zeros(k + 1)
    right = np.zeros(k + 1)
    for j in range(1, k + 1):
        left[j] = x - t[i + 1 - j]
        right[j] = t[i + j] - x
        saved = 0.0
        for r in range(j):
            tmp = w[r] / (right[r + 1] + left[j - r])
            w[r] = saved + right[r + 1] * tmp
            saved = left[j - r] * tmp
        w[j] = saved
    return i - k, w
