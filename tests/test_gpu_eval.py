"""-m gpu: the GPU evaluation-metric kernels (wgg_eval_*, csrc/eval.cu) against the CPU
restatement oracle/eval_oracle.py (pinned to the reference's evaluate_all_metrics by tests/test_oracle_golden.py).

Tolerances: distance matrices / assignment cost / jerk relative 1e-5 (fp32 kernels with fp64 accumulation vs float64);
correlations absolute 2e-4 (they are means over gestures of Pearson coefficients of float32 finite differences);
precision / recall: at most one gesture may change side (a distance within fp32 rounding of a k-NN radius)."""
import os

import numpy as np
import pytest
import torch

import wgg_b200 as wgg
from oracle import eval_oracle as E
from wgg_b200 import eval_metrics as M

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def fixture_pair(n, seed=0):
    fx = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "realistic_gestures.npz"))
    rng = np.random.default_rng(seed)
    real = fx["test_gesture"][:n].astype(np.float32)
    fake = fx["train_gesture"][:n].astype(np.float32) + (0.03 * rng.standard_normal((n, 128, 3))).astype(np.float32)
    fake[:, :, 2] = np.sort(np.clip(fake[:, :, 2], 0, 1), axis=1)
    return real, fake


@pytest.mark.parametrize("na,nb,d", [(1, 1, 1), (33, 70, 256), (300, 257, 256), (64, 64, 37)])
def test_cdist_and_kth(na, nb, d):
    rng = np.random.default_rng(na + nb)
    a = rng.standard_normal((na, d)).astype(np.float32)
    b = rng.standard_normal((nb, d)).astype(np.float32)
    ref = E.cdist_euclid(a, b)
    out = M.cdist(torch.from_numpy(a).to(DEV), torch.from_numpy(b).to(DEV))
    assert out.shape == (na, nb)
    assert np.abs(out.double().cpu().numpy() - ref).max() <= 1e-5 * max(1.0, ref.max())
    for k in (0, 3, 7):
        if k < nb:
            kth = M.row_kth(out, k).cpu().numpy()
            assert np.array_equal(kth, np.sort(out.cpu().numpy(), axis=1)[:, k])


@pytest.mark.parametrize("n", [8, 300])
def test_all_metrics_match_oracle(n):
    real, fake = fixture_pair(n)
    res = M.evaluate_all_metrics(real, fake, device=DEV)
    assert res["dtw_wasserstein"] == -1.0 and "fid" not in res
    ref = {"l2_wasserstein": E.l2_wasserstein(real, fake), "jerk_real": E.jerk(real), "jerk_fake": E.jerk(fake),
           "velocity_corr": E.velocity_corr(real, fake), "acceleration_corr": E.acceleration_corr(real, fake),
           "speed_profile_corr": E.speed_profile_corr(real, fake), "time_delta_corr": E.time_delta_corr(real, fake)}
    ref["precision"], ref["recall"] = E.precision_recall(real, fake, 3)
    for k in ("l2_wasserstein", "jerk_real", "jerk_fake"):
        assert abs(res[k] - ref[k]) <= 1e-5 * max(abs(ref[k]), 1e-6), (k, res[k], ref[k])
    for k in ("velocity_corr", "acceleration_corr", "speed_profile_corr", "time_delta_corr"):
        assert abs(res[k] - ref[k]) <= 2e-4, (k, res[k], ref[k])
    for k in ("precision", "recall"):
        assert abs(res[k] - ref[k]) <= 1.0 / n + 1e-9, (k, res[k], ref[k])
    # FID from supplied features = the oracle's Frechet distance
    rng = np.random.default_rng(1)
    rf, ff = rng.standard_normal((n + 40, 32)), rng.standard_normal((n + 40, 32)) * 1.1 + 0.1
    assert abs(M.fid_from_features(rf, ff, 32) - E.fid_from_features(rf, ff, 32)) <= 1e-9


def test_degenerate_gestures_are_skipped_like_the_reference():
    """Rows without variance (a constant time channel, a gesture that does not move) are left out of the correlation
    means exactly as evaluation.py:187,220,262,297 leaves them out; all-degenerate input gives 0."""
    real, fake = fixture_pair(6)
    real[0, :, 2] = 0.5                      # constant time: dt = 0 everywhere
    fake[1, :, :2] = 0.25                    # no movement
    res = M.dynamics_correlations(torch.from_numpy(real).to(DEV), torch.from_numpy(fake).to(DEV)).tolist()
    ref = [E.velocity_corr(real, fake), E.acceleration_corr(real, fake), E.speed_profile_corr(real, fake),
           E.time_delta_corr(real, fake)]
    for a, b in zip(res, ref):
        assert abs(a - b) <= 2e-4, (res, ref)
    z = np.zeros((3, 128, 3), np.float32)
    res0 = M.dynamics_correlations(torch.from_numpy(z).to(DEV), torch.from_numpy(z).to(DEV)).tolist()
    assert res0 == [0.0, 0.0, 0.0, 0.0]
    with pytest.raises(wgg._lib.WggError):
        M.row_kth(torch.zeros(4, 4, device=DEV), 8)
