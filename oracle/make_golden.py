"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):  python -m oracle.make_golden

For each case the reference's own ``WordGestureGANTrainer`` + ``train_epoch_with_grad_clip``
(src/gan/trainer.py, src/shared/utils.py:28-148) run ONE batch in float64 on CPU.  Noise is injected
by patching ``torch.randn`` / ``torch.randn_like`` in THIS process (the reference files are untouched)
so the 13 draws are explicit inputs; ``clip_grad_norm_`` is wrapped to record the un-clipped
gradients of each of the 12 optimiser steps.  The numpy oracle (oracle/wgg_oracle.py) is run on the
same inputs and must agree before the fixture is written.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import wgg_oracle as O  # noqa: E402
from oracle.ref_loader import load_reference  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

CASES = {
    # name: (model-config overrides, batch, seed)
    "tiny_temporal": (dict(seq_length=16, latent_dim=4, gen_hidden_dim=8, gen_num_layers=2,
                           enc_hidden_dims=(24, 12, 8, 6), disc_hidden_dims=(20, 12, 8, 6)), 3, 1),
    "tiny_mlp_time": (dict(seq_length=16, latent_dim=4, gen_hidden_dim=8, gen_num_layers=2,
                           enc_hidden_dims=(24, 12, 8, 6), disc_hidden_dims=(20, 12, 8, 6),
                           use_temporal_disc=False, prototype_has_time=True), 5, 2),
    "default": (dict(), 4, 3),
}


def sd_to_np(module, dtype=np.float64):
    return {k: v.detach().cpu().numpy().astype(dtype).copy() for k, v in module.state_dict().items()}


def run_reference_step(ref, mcfg_kwargs, B, seed):
    torch.set_default_dtype(torch.float64)
    try:
        mcfg = ref.config.ModelConfig(**mcfg_kwargs)
        tcfg = ref.config.TrainingConfig()
        ref.utils.seed_everything(42)
        trainer = ref.trainer.WordGestureGANTrainer(mcfg, tcfg, device="cpu")
        # round the initial state to fp32-representable values so fp32-stored fixtures are exact inputs
        with torch.no_grad():
            for mod in (trainer.generator, trainer.encoder, trainer.discriminator_1, trainer.discriminator_2):
                for t in list(mod.parameters()) + list(mod.buffers()):
                    t.copy_(t.float().double())
        init = {n: sd_to_np(getattr(trainer, m)) for n, m in
                (("G", "generator"), ("E", "encoder"), ("D1", "discriminator_1"), ("D2", "discriminator_2"))}
        g = torch.Generator().manual_seed(seed)
        T, Z = mcfg.seq_length, mcfg.latent_dim
        real = (torch.rand(B, T, 3, generator=g) * 2 - 1).float().double()
        proto = (torch.rand(B, T, 3, generator=g) * 2 - 1).float().double()
        n_draws = 2 * tcfg.n_critic + 3
        noise = [torch.randn(B, Z, generator=g).float().double() for _ in range(n_draws)]
        queue = [n.clone() for n in noise]

        orig_randn, orig_randn_like, orig_clip = torch.randn, torch.randn_like, torch.nn.utils.clip_grad_norm_
        grads_log = []

        def fake_randn(*a, **k):
            return queue.pop(0)

        def fake_randn_like(t, *a, **k):
            out = queue.pop(0)
            assert out.shape == t.shape
            return out

        def rec_clip(params, max_norm, *a, **k):
            params = list(params)
            grads_log.append([p.grad.detach().clone().numpy() for p in params])
            return orig_clip(params, max_norm, *a, **k)

        torch.randn, torch.randn_like, torch.nn.utils.clip_grad_norm_ = fake_randn, fake_randn_like, rec_clip
        cyc = {}
        o1, o2 = trainer.train_generator_step_cycle1, trainer.train_generator_step_cycle2

        def w1(p, r):
            out = o1(p, r)
            cyc["fake_cycle1"] = out[0].detach().numpy().copy()
            cyc.update(out[2])
            return out

        def w2(p, r):
            out = o2(p, r)
            cyc["fake_cycle2"] = out[0].detach().numpy().copy()
            cyc.update(out[2])
            return out

        trainer.train_generator_step_cycle1, trainer.train_generator_step_cycle2 = w1, w2
        try:
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                epoch = ref.utils.train_epoch_with_grad_clip(
                    trainer, [{"gesture": real, "prototype": proto}], 1.0, mcfg, tcfg, "cpu", scaler=None)
        finally:
            torch.randn, torch.randn_like, torch.nn.utils.clip_grad_norm_ = orig_randn, orig_randn_like, orig_clip
        assert not queue
        losses = dict(cyc)
        fakes = {k: losses.pop(k) for k in ("fake_cycle1", "fake_cycle2")}
        losses.update(epoch)
        post = {n: sd_to_np(getattr(trainer, m)) for n, m in
                (("G", "generator"), ("E", "encoder"), ("D1", "discriminator_1"), ("D2", "discriminator_2"))}
        adam = {}
        for n, opt, mod in (("G", trainer.optimizer_G, trainer.generator), ("E", trainer.optimizer_E, trainer.encoder),
                            ("D1", trainer.optimizer_D1, trainer.discriminator_1),
                            ("D2", trainer.optimizer_D2, trainer.discriminator_2)):
            names = [k for k, _ in mod.named_parameters()]
            st = opt.state_dict()["state"]
            adam[n] = {"names": names,
                       "m": {names[i]: st[i]["exp_avg"].numpy().copy() for i in st},
                       "v": {names[i]: st[i]["exp_avg_sq"].numpy().copy() for i in st},
                       "step": int(st[0]["step"])}
        # map recorded grads to names: order of clip calls = D1,D2 x n_critic, then G, E
        rec = {}
        idx = 0
        for it in range(tcfg.n_critic):
            for which, mod in (("D1", trainer.discriminator_1), ("D2", trainer.discriminator_2)):
                names = [k for k, _ in mod.named_parameters()]
                rec[f"{which}_grads_{it}"] = dict(zip(names, grads_log[idx]))
                idx += 1
        rec["G_grads"] = dict(zip([k for k, _ in trainer.generator.named_parameters()], grads_log[idx]))
        rec["E_grads"] = dict(zip([k for k, _ in trainer.encoder.named_parameters()], grads_log[idx + 1]))
        return dict(mcfg=mcfg, tcfg=tcfg, init=init, real=real.numpy(), proto=proto.numpy(),
                    noise=[n.numpy() for n in noise], losses=losses, fakes=fakes, post=post, adam=adam, rec=rec)
    finally:
        torch.set_default_dtype(torch.float32)


def run_oracle_step(r):
    mc = r["mcfg"]
    cfg = O.ModelCfg(seq_length=mc.seq_length, input_dim=mc.input_dim, latent_dim=mc.latent_dim,
                     gen_hidden_dim=mc.gen_hidden_dim, gen_num_layers=mc.gen_num_layers,
                     disc_hidden_dims=tuple(mc.disc_hidden_dims), use_temporal_disc=mc.use_temporal_disc,
                     prototype_has_time=mc.prototype_has_time, enc_hidden_dims=tuple(mc.enc_hidden_dims))
    tc = O.TrainCfg()
    s = O.GanState(*({k: v.copy() for k, v in r["init"][n].items()} for n in ("G", "E", "D1", "D2")))
    s.init_opt()
    rec = {}
    losses = O.train_batch(s, cfg, tc, r["real"], r["proto"], r["noise"], 1.0, None, rec)
    return s, losses, rec


def relerr(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    d = np.sqrt(((a - b) ** 2).sum())
    n = np.sqrt((b ** 2).sum())
    return d / n if n > 0 else d


def compare(r, s, losses, rec, tol=1e-9):
    worst = 0.0
    for k, v in r["losses"].items():
        e = abs(losses[k] - v) / max(abs(v), 1e-12)
        worst = max(worst, e)
        assert e < 1e-8, (k, losses[k], v)
    for k in ("fake_cycle1", "fake_cycle2"):
        e = relerr(rec[k], r["fakes"][k])
        worst = max(worst, e)
        assert e < tol, (k, e)
    for gk, gd in r["rec"].items():
        for name, g in gd.items():
            e = relerr(rec[gk][name], g)
            worst = max(worst, e)
            assert e < 1e-7, (gk, name, e)
    for n in ("G", "E", "D1", "D2"):
        sp = getattr(s, n)
        for name, v in r["post"][n].items():
            e = relerr(sp[name], v)
            worst = max(worst, e)
            assert e < 1e-7, ("post", n, name, e)
        for name, v in r["adam"][n]["m"].items():
            e = relerr(s.opt[n]["m"][name], v)
            assert e < 1e-6, ("adam m", n, name, e)
        for name, v in r["adam"][n]["v"].items():
            e = relerr(s.opt[n]["v"][name], v)
            assert e < 1e-6, ("adam v", n, name, e)
        assert s.opt[n]["step"] == r["adam"][n]["step"]
    return worst


def subsample(a, n=48):
    flat = np.asarray(a, np.float64).ravel()
    idx = np.linspace(0, flat.size - 1, min(n, flat.size)).astype(np.int64)
    return idx, flat[idx]


def summarise(a):
    a = np.asarray(a, np.float64)
    idx, vals = subsample(a)
    return np.concatenate([[a.sum(), np.abs(a).sum(), np.sqrt((a * a).sum())], vals])


def pack(name, r, full: bool):
    """Fixture layout: ``init/*`` full fp32 (exact: the run started from fp32-representable values);
    everything else either ``full/<kind>/...`` (fp64) or ``sum/<kind>/...`` = [sum, abs-sum, l2, 48 strided
    samples] (see ``summarise``).  Full arrays are kept for G/E and for D1's first critic step to bound size."""
    out = {}
    mc = r["mcfg"]
    out["cfg/seq_length"] = mc.seq_length
    out["cfg/latent_dim"] = mc.latent_dim
    out["cfg/gen_hidden_dim"] = mc.gen_hidden_dim
    out["cfg/gen_num_layers"] = mc.gen_num_layers
    out["cfg/enc_hidden_dims"] = np.array(mc.enc_hidden_dims)
    out["cfg/disc_hidden_dims"] = np.array(mc.disc_hidden_dims)
    out["cfg/use_temporal_disc"] = int(mc.use_temporal_disc)
    out["cfg/prototype_has_time"] = int(mc.prototype_has_time)
    out["real"] = r["real"]
    out["proto"] = r["proto"]
    out["noise"] = np.stack(r["noise"])
    for k, v in r["losses"].items():
        out[f"loss/{k}"] = np.float64(v)
    for k, v in r["fakes"].items():
        out[k] = v

    def put(kind, n, k, v, keep_full):
        if keep_full:
            out[f"full/{kind}/{n}/{k}"] = np.asarray(v, np.float64)
        else:
            out[f"sum/{kind}/{n}/{k}"] = summarise(v)

    for n in ("G", "E", "D1", "D2"):
        out[f"order/{n}"] = np.array(list(r["init"][n].keys()))
        out[f"porder/{n}"] = np.array(r["adam"][n]["names"])
        for k, v in r["init"][n].items():
            assert np.array_equal(v.astype(np.float32).astype(np.float64), v)
            out[f"init/{n}/{k}"] = v.astype(np.float32)
        ge = full and n in ("G", "E")
        for k, v in r["post"][n].items():
            put("post", n, k, v, ge or (full and n == "D1"))
        for k, v in r["adam"][n]["m"].items():
            put("adam_m", n, k, v, ge)
        for k, v in r["adam"][n]["v"].items():
            put("adam_v", n, k, v, ge)
        out[f"adam_step/{n}"] = r["adam"][n]["step"]
    for gk, gd in r["rec"].items():
        keep = full and gk in ("G_grads", "E_grads", "D1_grads_0")
        for k, v in gd.items():
            put("grad", gk, k, v, keep)
    path = os.path.join(GOLDEN_DIR, f"step_{name}.npz")
    np.savez_compressed(path, **out)
    return path


def write_init_fixture(ref):
    """Summaries of the reference's seed-42 fp32 initial state (default config), for the init-parity test."""
    ref.utils.seed_everything(42)
    trainer = ref.trainer.WordGestureGANTrainer(device="cpu")
    out = {}
    for n, m in (("G", "generator"), ("E", "encoder"), ("D1", "discriminator_1"), ("D2", "discriminator_2")):
        sd = getattr(trainer, m).state_dict()
        out[f"order/{n}"] = np.array(list(sd.keys()))
        for k, v in sd.items():
            out[f"{n}/{k}"] = summarise(v.numpy())
    path = os.path.join(GOLDEN_DIR, "init_seed42.npz")
    np.savez_compressed(path, **out)
    print(f"[golden] wrote {path}")


def main():
    ref = load_reference()
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    write_init_fixture(ref)
    for name, (kw, B, seed) in CASES.items():
        r = run_reference_step(ref, kw, B, seed)
        full = name != "default"
        s, losses, rec = run_oracle_step(r)
        worst = compare(r, s, losses, rec)
        path = pack(name, r, full)
        print(f"[golden] {name}: oracle == reference (worst rel err {worst:.2e}); wrote {path} "
              f"({os.path.getsize(path) / 1e6:.2f} MB)")


if __name__ == "__main__":
    main()
