"""CPU: the numpy oracle reproduces the reference-generated golden fixtures (full training batch:
all 11 losses, the un-clipped gradients of all 12 optimiser steps, post-step parameters, spectral-norm
buffers and Adam moments).  This pins oracle/wgg_oracle.py to the reference wherever the tests run."""
import numpy as np
import pytest

from golden_util import CASES, LOSS_KEYS, MODS, Golden, oracle_cfg
from oracle import wgg_oracle as O


@pytest.mark.parametrize("case", CASES)
def test_oracle_train_batch_matches_reference(case):
    g = Golden(case)
    cfg = oracle_cfg(g)
    tc = O.TrainCfg()
    s = O.GanState(*(g.init_state(m) for m in MODS))
    s.init_opt()
    real, proto, noise = g.inputs()
    rec = {}
    losses = O.train_batch(s, cfg, tc, real, proto, noise, 1.0, None, rec)
    for k in LOSS_KEYS:
        assert abs(losses[k] - g.loss(k)) <= 1e-9 * max(1.0, abs(g.loss(k))), k
    for k in ("fake_cycle1", "fake_cycle2"):
        assert np.abs(rec[k] - g.z[k]).max() < 1e-10
    for grp in [f"{d}_grads_{i}" for i in range(tc.n_critic) for d in ("D1", "D2")] + ["G_grads", "E_grads"]:
        for name, val in rec[grp].items():
            g.check("grad", grp, name, val, 1e-7, "oracle")
    for m in MODS:
        st = getattr(s, m)
        for name, val in st.items():
            g.check("post", m, name, val, 1e-7, "oracle")
        for name in g.param_order(m):
            g.check("adam_m", m, name, s.opt[m]["m"][name], 1e-6, "oracle")
            g.check("adam_v", m, name, s.opt[m]["v"][name], 1e-6, "oracle")
        assert s.opt[m]["step"] == int(g.z[f"adam_step/{m}"])


def test_noise_draw_count():
    assert O.n_noise_draws(O.TrainCfg()) == 13


def test_feature_matching_double_normalisation():
    # losses.py:88-93: mean over all elements AND a division by the per-sample feature count
    r = [np.zeros((2, 4))]
    f = [np.ones((2, 4))]
    loss, grads = O.feature_matching(r, f)
    assert abs(loss - 1.0 / 4.0) < 1e-15
    assert np.allclose(grads[0], 1.0 / (8 * 4))


def test_spectral_norm_gradient_formula():
    rng = np.random.default_rng(0)
    W = rng.standard_normal((5, 7))
    p = {"l.weight_orig": W, "l.weight_u": O._normalize(rng.standard_normal(5)),
         "l.weight_v": O._normalize(rng.standard_normal(7))}
    _, sigma, u, v = O.sn_effective_weight(p, "l", True)
    G = rng.standard_normal((5, 7))
    an = O.sn_weight_grad(W, G, sigma, u, v)
    eps = 1e-6
    num = np.zeros_like(W)
    for i in range(5):
        for j in range(7):
            Wp = W.copy(); Wp[i, j] += eps
            Wm = W.copy(); Wm[i, j] -= eps
            fp = (G * (Wp / float(u @ (Wp @ v)))).sum()
            fm = (G * (Wm / float(u @ (Wm @ v)))).sum()
            num[i, j] = (fp - fm) / (2 * eps)
    assert np.abs(an - num).max() < 1e-6


@pytest.mark.parametrize("case", ["tiny_temporal", "tiny_mlp_time"])
def test_torch_port_matches_reference_golden(case):
    """oracle/torch_port.py (the CPU arm bench.py times) reproduces the reference's losses and post-step state."""
    import torch
    from oracle import torch_port
    g = Golden(case)
    cfg = oracle_cfg(g)
    tp = torch_port.TorchPortTrainer(seed=0, cfg=cfg, dtype=torch.float64)
    tp.load_state({m: g.init_state(m) for m in MODS})
    real, proto, noise = g.inputs()
    out = tp.train_batch(torch.from_numpy(real), torch.from_numpy(proto), noise)
    for k in LOSS_KEYS:
        assert abs(out[k] - g.loss(k)) <= 1e-9 * max(1.0, abs(g.loss(k))), k
    st = tp.state()
    for m in MODS:
        for name, val in st[m].items():
            g.check("post", m, name, val, 1e-7, "torch_port")
