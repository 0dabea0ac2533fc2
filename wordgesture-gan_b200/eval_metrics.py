"""Evaluation metrics of generated gestures on the GPU.

Mirror of the reference's ``evaluate_all_metrics`` (src/gan/evaluation.py:297-500): same argument meaning, same
result keys.  The O(n^2 d) distance matrices, the k-NN manifold precision / recall, the Savitzky-Golay jerk and the four
time-aware dynamics correlations run as CUDA kernels through the C ABI (csrc/eval.cu); what stays on the host is what
is inherently sequential and tiny: the optimal assignment on the distance matrix (scipy's linear_sum_assignment, as in
the reference, evaluation.py:336) and the 32 x 32 matrix square root of the Frechet distance (evaluation.py:460).
The FID auto-encoder is an evaluation-only torch model of the reference (src/gan/models.py:356-450, out of scope of
this package): pass its features (``real_features`` / ``fake_features``) or an object with ``.encode`` to get ``fid``.
DTW-Wasserstein (fastdtw, evaluation.py:341-361) is not implemented: ``dtw_wasserstein`` is -1.0, the value the
reference reports under ``skip_dtw=True``.
"""
from __future__ import annotations

from math import factorial
from typing import Dict, Optional

import numpy as np
import torch

from . import _lib


def savgol_operator(n: int, window: int, poly: int, deriv: int) -> np.ndarray:
    """(n, n) matrix of the linear map x -> savgol_filter(x, window, poly, deriv=deriv, delta=1, mode='interp'):
    interior rows carry the least-squares derivative stencil; the first / last window // 2 rows evaluate the derivative
    of the polynomial fitted to the first / last ``window`` samples (scipy's edge handling)."""
    half = window // 2
    S = np.zeros((n, n))
    if deriv > poly:
        return S
    pos = np.arange(-half, half + 1, dtype=np.float64)
    P = np.linalg.pinv(np.vander(pos, poly + 1, increasing=True))
    for i in range(half, n - half):
        S[i, i - half:i + half + 1] = factorial(deriv) * P[deriv]
    Pl = np.linalg.pinv(np.vander(np.arange(window, dtype=np.float64), poly + 1, increasing=True))
    for i in range(half):
        for row_i, cols, x0 in ((i, slice(0, window), float(i)), (n - half + i, slice(n - window, n), float(window - half + i))):
            row = np.zeros(window)
            for p in range(deriv, poly + 1):
                row += (factorial(p) / factorial(p - deriv)) * x0 ** (p - deriv) * Pl[p]
            S[row_i, cols] = row
    return S


def _dev(t, device) -> torch.Tensor:
    if isinstance(t, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(t, dtype=np.float32))
    return t.to(device=device, dtype=torch.float32).contiguous()


def cdist(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """Euclidean distance matrix (na, nb) of the rows of a (na, d) and b (nb, d)."""
    c = _lib.ctx(a.device)
    out = torch.empty(a.shape[0], b.shape[0], dtype=torch.float32, device=a.device)
    _lib.check(_lib.lib().wgg_eval_cdist(c, _lib.ptr(a), a.shape[0], _lib.ptr(b), b.shape[0], a.shape[1], _lib.ptr(out),
                                         _lib.stream(a.device)), c)
    return out


def row_kth(m: torch.Tensor, k: int) -> torch.Tensor:
    """np.sort(m, axis=1)[:, k] for k < 8."""
    c = _lib.ctx(m.device)
    out = torch.empty(m.shape[0], dtype=torch.float32, device=m.device)
    _lib.check(_lib.lib().wgg_eval_row_kth(c, _lib.ptr(m), m.shape[0], m.shape[1], k, _lib.ptr(out), _lib.stream(m.device)), c)
    return out


def precision_recall(real_fake: torch.Tensor, real_radii: torch.Tensor, fake_radii: torch.Tensor):
    c = _lib.ctx(real_fake.device)
    out = torch.empty(2, dtype=torch.float32, device=real_fake.device)
    ws = torch.empty(2, dtype=torch.float32, device=real_fake.device)
    _lib.check(_lib.lib().wgg_eval_precision_recall(c, _lib.ptr(real_fake), real_fake.shape[0], real_fake.shape[1],
                                                    _lib.ptr(real_radii), _lib.ptr(fake_radii), _lib.ptr(out), _lib.ptr(ws),
                                                    _lib.stream(real_fake.device)), c)
    return out


def jerk(g: torch.Tensor, window: int = 21, poly: int = 3) -> torch.Tensor:
    """0-dim tensor: mean Savitzky-Golay jerk magnitude (evaluation.py:364-374); 0 for sequences shorter than the window."""
    n, T, C = g.shape
    if T < window:
        return torch.zeros((), dtype=torch.float32, device=g.device)
    S = _dev(savgol_operator(T, window, poly, 3), g.device)
    c = _lib.ctx(g.device)
    out = torch.empty(1, dtype=torch.float32, device=g.device)
    ws = torch.empty(n, dtype=torch.float32, device=g.device)
    _lib.check(_lib.lib().wgg_eval_jerk(c, _lib.ptr(g), n, T, C, _lib.ptr(S), _lib.ptr(out), _lib.ptr(ws), _lib.stream(g.device)), c)
    return out[0]


def dynamics_correlations(real: torch.Tensor, fake: torch.Tensor) -> torch.Tensor:
    """(4,) tensor: velocity, acceleration, speed-profile and time-delta correlations (evaluation.py:162-305)."""
    n, T, C = real.shape
    c = _lib.ctx(real.device)
    out = torch.empty(4, dtype=torch.float32, device=real.device)
    ws = torch.empty(8 * n, dtype=torch.float32, device=real.device)
    _lib.check(_lib.lib().wgg_eval_dynamics(c, _lib.ptr(real), _lib.ptr(fake), n, T, C, _lib.ptr(out), _lib.ptr(ws),
                                            _lib.stream(real.device)), c)
    return out


def fid_from_features(real_features, fake_features, dim: int) -> float:
    """Frechet distance of two feature clouds (evaluation.py:456-464); 32 x 32 host linear algebra."""
    from scipy.linalg import sqrtm
    rf = np.asarray(real_features, np.float64)
    ff = np.asarray(fake_features, np.float64)
    mu_r, mu_f = rf.mean(axis=0), ff.mean(axis=0)
    cr = np.cov(rf, rowvar=False) + np.eye(dim) * 1e-6
    cf = np.cov(ff, rowvar=False) + np.eye(dim) * 1e-6
    covmean = sqrtm(cr @ cf).real
    return float(((mu_r - mu_f) ** 2).sum() + np.trace(cr + cf - 2 * covmean))


def evaluate_all_metrics(real_gestures, fake_gestures, device="cuda", precision_recall_k: int = 3, savgol_window: int = 21,
                         savgol_poly_order: int = 3, fid_hidden_dim: int = 32, autoencoder=None,
                         real_features: Optional[np.ndarray] = None, fake_features: Optional[np.ndarray] = None
                         ) -> Dict[str, float]:
    """GPU counterpart of src/gan/evaluation.py:297 for equally many real and fake gestures (n, T, 3).  Keyword
    defaults are EvaluationConfig's (src/shared/config.py:69-87).  Returns the reference's keys (``dtw_wasserstein``
    = -1.0; ``fid`` only when features or an encoder are supplied)."""
    from scipy.optimize import linear_sum_assignment
    dev = torch.device(device)
    real = _dev(real_gestures, dev)
    fake = _dev(fake_gestures, dev)
    if real.shape != fake.shape or real.dim() != 3 or real.shape[2] < 3:
        raise ValueError(f"real and fake must both be (n, T, 3), got {tuple(real.shape)} and {tuple(fake.shape)}")
    n = real.shape[0]
    res: Dict[str, float] = {}
    real_xy = real[:, :, :2].reshape(n, -1).contiguous()
    fake_xy = fake[:, :, :2].reshape(n, -1).contiguous()
    rf = cdist(real_xy, fake_xy)
    cost = rf.double().cpu().numpy()
    r, c = linear_sum_assignment(cost)              # sequential O(n^3): host, as in the reference (evaluation.py:336)
    res["l2_wasserstein"] = float(cost[r, c].mean())
    res["dtw_wasserstein"] = -1.0
    res["jerk_real"] = float(jerk(real, savgol_window, savgol_poly_order))
    res["jerk_fake"] = float(jerk(fake, savgol_window, savgol_poly_order))
    dyn = dynamics_correlations(real, fake).tolist()
    res["velocity_corr"], res["acceleration_corr"], res["speed_profile_corr"], res["time_delta_corr"] = dyn
    if autoencoder is not None and (real_features is None or fake_features is None):
        with torch.no_grad():
            real_features = autoencoder.encode(real).cpu().numpy()
            fake_features = autoencoder.encode(fake).cpu().numpy()
    if real_features is not None and fake_features is not None:
        res["fid"] = fid_from_features(real_features, fake_features, fid_hidden_dim)
    real_radii = row_kth(cdist(real_xy, real_xy), precision_recall_k)
    fake_radii = row_kth(cdist(fake_xy, fake_xy), precision_recall_k)
    pr = precision_recall(rf, real_radii, fake_radii).tolist()
    res["precision"], res["recall"] = pr
    return res
