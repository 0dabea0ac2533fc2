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


@pytest.mark.parametrize("case", ["tiny_temporal", "tiny_mlp_time", "default"])
def test_torch_port_matches_reference_golden(case):
    """oracle/torch_port.py (the CPU arm bench.py times, and the fp64 checker of the large-batch GPU parity tests)
    reproduces the reference's losses, the un-clipped gradients of all 12 optimiser steps, both generator-side fake
    gestures and the post-step state - on the default model too."""
    import torch
    from oracle import torch_port
    g = Golden(case)
    cfg = oracle_cfg(g)
    tp = torch_port.TorchPortTrainer(seed=0, cfg=cfg, dtype=torch.float64)
    tp.load_state({m: g.init_state(m) for m in MODS})
    real, proto, noise = g.inputs()
    rec, fakes = {}, {}
    out = tp.train_batch(torch.from_numpy(real), torch.from_numpy(proto), noise, record=rec, fakes=fakes)
    for k in LOSS_KEYS:
        assert abs(out[k] - g.loss(k)) <= 1e-9 * max(1.0, abs(g.loss(k))), k
    assert len(rec) == 2 * O.TrainCfg().n_critic + 2
    for tag, d in rec.items():
        for name, val in d.items():
            g.check("grad", tag, name, val, 1e-7, "torch_port")
    assert np.abs(fakes["fake1"] - g.z["fake_cycle1"]).max() <= 1e-9
    assert np.abs(fakes["fake2"] - g.z["fake_cycle2"]).max() <= 1e-9
    st = tp.state()
    for m in MODS:
        for name, val in st[m].items():
            g.check("post", m, name, val, 1e-7, "torch_port")


@pytest.mark.parametrize("case", ["tiny_temporal", "tiny_mlp_time", "default"])
def test_torch_port_cycles_match_reference_direct_calls(case):
    """The CPU restatement of train_generator_step_cycle1/2 (trainer.py:84-193) against fixtures produced by calling
    those methods of the unmodified reference directly (oracle/make_golden_cycles.py): returned fake gesture, loss
    dict, total, generator / encoder gradients of each cycle on its own, and the discriminator's u/v buffers after
    its three calls."""
    import torch
    from golden_util import CycleGolden, rel_l2
    from oracle import torch_port
    g = CycleGolden(case)
    cfg = O.ModelCfg(**g.cfg_kwargs())
    tp = torch_port.TorchPortTrainer(seed=0, cfg=cfg, dtype=torch.float64)
    tp.load_state({m: g.init_state(m) for m in MODS})
    real, proto, z, eps_rec, eps = g.inputs()
    real, proto = torch.from_numpy(real), torch.from_numpy(proto)
    for cyc, call, disc in ((1, lambda: tp.cycle1(proto, real, z, eps_rec), tp.D1), (2, lambda: tp.cycle2(proto, real, eps), tp.D2)):
        tp.opt["G"].zero_grad()
        tp.opt["E"].zero_grad()
        fake, total, d = call()
        total.backward()
        assert np.abs(fake.detach().numpy() - g.z[f"c{cyc}/fake"]).max() <= 1e-10
        assert abs(total.item() - float(g.z[f"c{cyc}/total"])) <= 1e-10
        ref_d = g.losses(cyc)
        assert set(d) == set(ref_d)
        for k, v in ref_d.items():
            assert abs(d[k] - v) <= 1e-10 * max(1.0, abs(v)), k
        for mod, net in (("G", tp.G), ("E", tp.E)):
            ref_g = g.grads(cyc, mod)
            for k, p in net.named_parameters():
                mine = np.zeros(tuple(p.shape)) if p.grad is None else p.grad.numpy()
                if np.abs(ref_g[k]).max() == 0:
                    assert np.abs(mine).max() == 0, (cyc, mod, k)  # cycle 1 carries no encoder gradient (trainer.py:116-119)
                else:
                    assert rel_l2(mine, ref_g[k]) <= 1e-8, (cyc, mod, k)
        sd = disc.state_dict()
        for k, v in g.uv(cyc).items():
            assert np.abs(sd[k].numpy() - v).max() <= 1e-10, k


def test_eval_oracle_matches_reference(tmp_path):
    """oracle/eval_oracle.py (the checker of the GPU metric kernels) against the reference's own
    evaluate_all_metrics (src/gan/evaluation.py:297-500) and its time-aware correlation functions, on a slice of the
    realistic fixture.  Needs the reference's files (/root/reference, or oracle/_ref on the GPU box)."""
    import torch
    from oracle import eval_oracle as E
    from oracle.ref_loader import load_reference, reference_available
    if not reference_available():
        pytest.skip("reference sources not reachable")
    ref = load_reference()
    ev = ref.evaluation
    import os
    fx = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "realistic_gestures.npz"))
    n = 64
    real = fx["test_gesture"][:n].astype(np.float64)
    rng = np.random.default_rng(3)
    fake = fx["test_gesture"][n:2 * n].astype(np.float64) + 0.02 * rng.standard_normal((n, 128, 3))
    fake[:, :, 2] = np.sort(np.clip(fake[:, :, 2], 0, 1), axis=1)
    assert abs(E.velocity_corr(real, fake) - ev.time_aware_velocity_correlation(real, fake)) <= 1e-12
    assert abs(E.acceleration_corr(real, fake) - ev.time_aware_acceleration_correlation(real, fake)) <= 1e-12
    assert abs(E.speed_profile_corr(real, fake) - ev.speed_profile_correlation(real, fake)) <= 1e-12
    assert abs(E.time_delta_corr(real, fake) - ev.time_delta_correlation(real, fake)) <= 1e-12
    ecfg = ref.config.EvaluationConfig(fid_autoencoder_epochs=1)
    old = ev._get_ae_cache_path
    ev._get_ae_cache_path = lambda train_data, eval_config: tmp_path / "ae.pt"
    try:
        torch.manual_seed(0)
        res = ev.evaluate_all_metrics(real.astype(np.float32), fake.astype(np.float32), eval_config=ecfg, device="cpu",
                                      skip_dtw=True)
    finally:
        ev._get_ae_cache_path = old
    r32, f32_ = real.astype(np.float32), fake.astype(np.float32)
    assert abs(E.l2_wasserstein(r32, f32_) - res["l2_wasserstein"]) <= 1e-9
    assert abs(E.jerk(r32) - res["jerk_real"]) <= 1e-9 * max(1.0, res["jerk_real"])
    assert abs(E.jerk(f32_) - res["jerk_fake"]) <= 1e-9 * max(1.0, res["jerk_fake"])
    prec, rec = E.precision_recall(r32, f32_, ecfg.precision_recall_k)
    assert prec == res["precision"] and rec == res["recall"]
    ae = res["_cached_real"]["autoencoder"]
    with torch.no_grad():
        ff = ae.encode(torch.from_numpy(f32_)).numpy()
    fid = E.fid_from_features(res["_cached_real"]["real_features"], ff, ecfg.fid_hidden_dim)
    assert abs(fid - res["fid"]) <= 1e-8 * max(1.0, abs(res["fid"]))


def test_step_restructurings_are_exact():
    """The product path restructures the step without changing its numbers (train_step.py / gan_trainer.py):
    (a) ONE generator call on the stacked critic-phase batch instead of 2*n_critic calls,
    (b) ONE encoder pass on the real batch + n re-parameterisations (mu / log_var do not depend on eps),
    (c) cycle-1 and cycle-2 generator passes as one stacked call.
    Checked here on the CPU restatement of the reference's modules (fp64): no layer mixes samples, so the stacked
    calls reproduce the separate calls row for row."""
    import torch
    from oracle import torch_port
    g = Golden("tiny_temporal")
    ocfg = oracle_cfg(g)
    tp = torch_port.TorchPortTrainer(seed=0, cfg=ocfg, tc=O.TrainCfg(), dtype=torch.float64)
    tp.load_state({m: g.init_state(m) for m in MODS})
    gen = torch.Generator().manual_seed(3)
    B, n = 6, 3
    real = torch.rand(B, ocfg.seq_length, ocfg.input_dim, generator=gen, dtype=torch.float64) * 2 - 1
    proto = torch.rand(B, ocfg.seq_length, ocfg.input_dim, generator=gen, dtype=torch.float64) * 2 - 1
    zs = [torch.randn(B, ocfg.latent_dim, generator=gen, dtype=torch.float64) for _ in range(n)]
    epss = [torch.randn(B, ocfg.latent_dim, generator=gen, dtype=torch.float64) for _ in range(n)]
    with torch.no_grad():
        # (b) encoder: z_i = mu + eps_i * exp(0.5 log_var) with the mu / log_var of a single pass
        z0, mu, lv = tp.E(real, epss[0])
        for e in epss:
            zi, mui, lvi = tp.E(real, e)
            assert torch.equal(mui, mu) and torch.equal(lvi, lv)
            assert torch.allclose(zi, torch.addcmul(mu, e, torch.exp(0.5 * lv)), rtol=0, atol=1e-15)
        # (a) + (c) generator on stacked batches
        z_enc = [tp.E(real, e)[0] for e in epss]
        stacked = tp.G(proto.repeat(2 * n, 1, 1), torch.cat(zs + z_enc, 0))
        for i, z in enumerate(zs + z_enc):
            single = tp.G(proto, z)
            assert torch.allclose(stacked[i * B:(i + 1) * B], single, rtol=0, atol=1e-13), i
