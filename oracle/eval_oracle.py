"""CPU restatement (numpy / scipy, float64) of the reference's evaluation metrics, src/gan/evaluation.py.

TEST INFRASTRUCTURE (see oracle/__init__.py): the checker of the GPU metric kernels.  Pinned to
the reference's own functions by tests/test_oracle_golden.py::test_eval_oracle_matches_reference (runs wherever the
reference's files are reachable - /root/reference or the vendored oracle/_ref).  Each function cites the lines it restates.
"""
from __future__ import annotations

import numpy as np


def cdist_euclid(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """scipy.spatial.distance.cdist(a, b, 'euclidean') as used at evaluation.py:335, 474-476."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    out = np.empty((a.shape[0], b.shape[0]))
    for i in range(a.shape[0]):
        d = b - a[i]
        out[i] = np.sqrt((d * d).sum(axis=1))
    return out


def flat_xy(g: np.ndarray) -> np.ndarray:
    """evaluation.py:331,334: the (x, y) channels flattened per gesture."""
    return np.asarray(g)[:, :, :2].reshape(len(g), -1)


def l2_wasserstein(real: np.ndarray, fake: np.ndarray) -> float:
    """evaluation.py:333-337: mean cost of the optimal one-to-one assignment on the Euclidean cost matrix."""
    from scipy.optimize import linear_sum_assignment
    d = cdist_euclid(flat_xy(real), flat_xy(fake))
    r, c = linear_sum_assignment(d)
    return float(d[r, c].mean())


def savgol_operator(n: int, window: int, poly: int, deriv: int) -> np.ndarray:
    """The linear map x -> scipy.signal.savgol_filter(x, window, poly, deriv=deriv) (delta = 1, mode='interp') as an
    (n, n) matrix: interior rows hold the least-squares derivative stencil, the first / last window // 2 rows evaluate
    the derivative of the polynomial fitted to the first / last `window` samples (scipy's 'interp' edge handling)."""
    from math import factorial
    half = window // 2
    S = np.zeros((n, n))
    pos = np.arange(-half, half + 1, dtype=np.float64)
    V = np.vander(pos, poly + 1, increasing=True)           # V[j, p] = pos_j ** p
    P = np.linalg.pinv(V)                                    # coefficients = P @ samples
    stencil = factorial(deriv) * P[deriv] if deriv <= poly else np.zeros(window)
    for i in range(half, n - half):
        S[i, i - half:i + half + 1] = stencil
    # edges: polynomial through the first / last window, derivative evaluated at each edge position
    pos_l = np.arange(window, dtype=np.float64)
    Pl = np.linalg.pinv(np.vander(pos_l, poly + 1, increasing=True))
    for i in range(half):
        for edge_i, cols in ((i, slice(0, window)), (n - half + i, slice(n - window, n))):
            x0 = float(i) if edge_i < half else float(window - half + i)
            row = np.zeros(window)
            for p in range(deriv, poly + 1):
                row += (factorial(p) / factorial(p - deriv)) * x0 ** (p - deriv) * Pl[p]
            S[edge_i, cols] = row
    return S


def jerk(gestures: np.ndarray, window: int = 21, poly: int = 3) -> float:
    """evaluation.py:364-374: mean over gestures of mean_t sqrt(x'''^2 + y'''^2), third derivatives by
    savgol_filter(window, poly, deriv=3); 0 for sequences shorter than the window."""
    g = np.asarray(gestures, np.float64)
    n = g.shape[1]
    if n < window:
        return 0.0
    S = savgol_operator(n, window, poly, 3)
    d3x = g[:, :, 0] @ S.T
    d3y = g[:, :, 1] @ S.T
    return float(np.sqrt(d3x ** 2 + d3y ** 2).mean(axis=1).mean())


def _velocity(g):
    """evaluation.py:55-89 compute_time_aware_velocity."""
    xy, t = g[:, :, :2], g[:, :, 2]
    dxy, dt = np.diff(xy, axis=1), np.diff(t, axis=1)
    t_mid = (t[:, :-1] + t[:, 1:]) / 2
    dt_safe = np.where(np.abs(dt) > 1e-10, dt, 1e-10 * np.sign(dt + 1e-20))
    return dxy / dt_safe[:, :, None], t_mid


def _acceleration(g):
    """evaluation.py:92-123 compute_time_aware_acceleration."""
    v, t_mid = _velocity(g)
    dv, dtm = np.diff(v, axis=1), np.diff(t_mid, axis=1)
    dtm_safe = np.where(np.abs(dtm) > 1e-10, dtm, 1e-10 * np.sign(dtm + 1e-20))
    return dv / dtm_safe[:, :, None]


def _mean_corr(a, b, lo_a, hi_a, lo_b, hi_b):
    """The per-gesture loop shared by evaluation.py:183-196, 216-228, 258-271, 293-303: Pearson correlation of the
    clipped rows, skipping rows without variance and NaN results; mean over the kept rows (0 if none)."""
    cors = []
    for i in range(len(a)):
        r, f = a[i], b[i]
        if len(r) > 1 and np.std(r) > 1e-10 and np.std(f) > 1e-10:
            rc = np.clip(r, lo_a(r), hi_a(r)) if lo_a else r
            fc = np.clip(f, lo_b(f), hi_b(f)) if lo_b else f
            with np.errstate(invalid="ignore", divide="ignore"):
                c = np.corrcoef(rc, fc)[0, 1]
            if not np.isnan(c):
                cors.append(c)
    return float(np.mean(cors)) if cors else 0.0


_p1 = lambda v: np.percentile(v, 1)
_p99 = lambda v: np.percentile(v, 99)
_zero = lambda v: 0


def velocity_corr(real, fake) -> float:
    """evaluation.py:162-198."""
    vr, _ = _velocity(np.asarray(real, np.float64))
    vf, _ = _velocity(np.asarray(fake, np.float64))
    return _mean_corr(vr.reshape(len(vr), -1), vf.reshape(len(vf), -1), _p1, _p99, _p1, _p99)


def acceleration_corr(real, fake) -> float:
    """evaluation.py:201-230."""
    ar = _acceleration(np.asarray(real, np.float64))
    af = _acceleration(np.asarray(fake, np.float64))
    return _mean_corr(ar.reshape(len(ar), -1), af.reshape(len(af), -1), _p1, _p99, _p1, _p99)


def speed_profile_corr(real, fake) -> float:
    """evaluation.py:233-273: |velocity| profiles, clipped to [0, 99th percentile]."""
    vr, _ = _velocity(np.asarray(real, np.float64))
    vf, _ = _velocity(np.asarray(fake, np.float64))
    return _mean_corr(np.linalg.norm(vr, axis=-1), np.linalg.norm(vf, axis=-1), _zero, _p99, _zero, _p99)


def time_delta_corr(real, fake) -> float:
    """evaluation.py:276-305: correlation of the raw time deltas (no clipping)."""
    dr = np.diff(np.asarray(real, np.float64)[:, :, 2], axis=1)
    df = np.diff(np.asarray(fake, np.float64)[:, :, 2], axis=1)
    return _mean_corr(dr, df, None, None, None, None)


def fid_from_features(real_feat, fake_feat, dim: int) -> float:
    """evaluation.py:456-464: Frechet distance of the two feature clouds (covariances regularised by 1e-6 I)."""
    from scipy.linalg import sqrtm
    rf, ff = np.asarray(real_feat, np.float64), np.asarray(fake_feat, np.float64)
    mu_r, mu_f = rf.mean(axis=0), ff.mean(axis=0)
    cr = np.cov(rf, rowvar=False) + np.eye(dim) * 1e-6
    cf = np.cov(ff, rowvar=False) + np.eye(dim) * 1e-6
    covmean = sqrtm(cr @ cf).real
    return float(((mu_r - mu_f) ** 2).sum() + np.trace(cr + cf - 2 * covmean))


def precision_recall(real, fake, k: int = 3):
    """evaluation.py:466-486: k-NN manifold precision / recall on the flattened (x, y) trajectories."""
    r, f = flat_xy(real), flat_xy(fake)
    real_radii = np.sort(cdist_euclid(r, r), axis=1)[:, k]
    fake_radii = np.sort(cdist_euclid(f, f), axis=1)[:, k]
    rf = cdist_euclid(r, f)
    n = len(r)
    prec = np.mean([np.any(rf[:, j] <= real_radii) for j in range(n)])
    rec = np.mean([np.any(rf[i, :] <= fake_radii) for i in range(n)])
    return float(prec), float(rec)
