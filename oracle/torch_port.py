"""CPU port of the training step on the reference's own library path (torch CPU kernels: oneDNN LSTM / conv,
ATen elementwise, torch.optim.Adam), used as the CPU arm that bench.py times (`cpu_baseline`, `--impl reference`).

TEST / BASELINE INFRASTRUCTURE - see oracle/__init__.py.  The reference itself is pure Python and lives only in
the build container, so this file restates its arithmetic (src/gan/models.py:52-353, src/gan/losses.py:43-175,
src/gan/trainer.py:102-183, src/shared/utils.py:62-135) compactly on torch.nn building blocks with autograd doing
the backward.  tests/test_oracle_golden.py::test_torch_port_* pins it to the reference-generated goldens.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn.utils import spectral_norm as sn

from .wgg_oracle import ModelCfg, TrainCfg


def _seq(*mods):
    return nn.Sequential(*mods)


class _G(nn.Module):
    def __init__(s, c: ModelCfg):
        super().__init__()
        s.pd = c.input_dim if c.prototype_has_time else 2
        s.lstm = nn.LSTM(s.pd + c.latent_dim, c.gen_hidden_dim, c.gen_num_layers, batch_first=True, bidirectional=True)
        s.output_layer = nn.Linear(2 * c.gen_hidden_dim, c.input_dim)

    def forward(s, proto, z):
        x = torch.cat([proto[..., :s.pd], z[:, None, :].expand(-1, proto.shape[1], -1)], -1)
        return torch.tanh(s.output_layer(s.lstm(x)[0]))


class _E(nn.Module):
    def __init__(s, c: ModelCfg):
        super().__init__()
        dims = [c.seq_length * c.input_dim, *c.enc_hidden_dims]
        layers = []
        for a, b in zip(dims, dims[1:]):
            layers += [nn.Linear(a, b), nn.LeakyReLU(0.2)]
        s.encoder = _seq(*layers)
        s.fc_mu = nn.Linear(dims[-1], c.latent_dim)
        s.fc_log_var = nn.Linear(dims[-1], c.latent_dim)

    def forward(s, x, eps):
        h = s.encoder(x.flatten(1))
        mu, lv = s.fc_mu(h), s.fc_log_var(h)
        return mu + eps * torch.exp(0.5 * lv), mu, lv


class _DT(nn.Module):
    def __init__(s, c: ModelCfg):
        super().__init__()
        act = lambda: nn.LeakyReLU(0.2)
        s.temporal_conv = _seq(sn(nn.Conv1d(c.input_dim, 64, 5, padding=2)), act(), sn(nn.Conv1d(64, 64, 5, padding=2)),
                               act(), sn(nn.Conv1d(64, 32, 3, padding=1)), act())
        s.mlp = _seq(sn(nn.Linear(256, 128)), act(), sn(nn.Linear(128, 64)), act())
        s.output_layer = sn(nn.Linear(64, 1))

    def feats(s, x):
        out, h = [], x.transpose(1, 2)
        for m in s.temporal_conv:
            h = m(h)
            if isinstance(m, nn.LeakyReLU):
                out.append(h.flatten(1))
        h = F.adaptive_avg_pool1d(h, 8).flatten(1)
        for m in s.mlp:
            h = m(h)
            if isinstance(m, nn.LeakyReLU):
                out.append(h)
        return out

    def forward(s, x):
        return s.output_layer(s.feats(x)[-1])


class _DM(nn.Module):
    def __init__(s, c: ModelCfg):
        super().__init__()
        dims = [c.seq_length * c.input_dim, *c.disc_hidden_dims]
        s.layers = nn.ModuleList(sn(nn.Linear(a, b)) for a, b in zip(dims, dims[1:]))
        s.output_layer = sn(nn.Linear(dims[-1], 1))

    def feats(s, x):
        out, h = [], x.flatten(1)
        for m in s.layers:
            h = F.leaky_relu(m(h), 0.2)
            out.append(h)
        return out

    def forward(s, x):
        return s.output_layer(s.feats(x)[-1])


def _fm(real_f, fake_f):
    return sum(F.l1_loss(f, r.detach()) / (r.numel() / r.shape[0]) for r, f in zip(real_f, fake_f)) / len(real_f)


class TorchPortTrainer:
    def __init__(self, seed: int = 42, cfg: ModelCfg = None, tc: TrainCfg = None, dtype=torch.float32, device="cpu"):
        """``device="cuda"`` runs the same restatement on the library kernels the reference reaches on a GPU (cuDNN
        LSTM / conv with TF32 allowed, fp32 cuBLAS - torch's defaults, which the reference does not change): the
        "reference on CUDA" bar of SURVEY.md 8(d)."""
        self.cfg, self.tc = cfg or ModelCfg(), tc or TrainCfg()
        self.device = torch.device(device)
        torch.manual_seed(seed)
        D = _DT if self.cfg.use_temporal_disc else _DM
        self.G, self.E, self.D1, self.D2 = _G(self.cfg), _E(self.cfg), D(self.cfg), D(self.cfg)
        self.mods = {"G": self.G, "E": self.E, "D1": self.D1, "D2": self.D2}
        for m in self.mods.values():
            m.to(dtype).to(self.device).train()
        self.dtype = dtype
        self.opt = {k: torch.optim.Adam(m.parameters(), lr=self.tc.learning_rate, betas=self.tc.betas)
                    for k, m in self.mods.items()}

    def load_state(self, states):
        for k, m in self.mods.items():
            m.load_state_dict({n: torch.as_tensor(v, dtype=self.dtype, device=self.device) for n, v in states[k].items()})

    def state(self):
        return {k: {n: v.detach().double().cpu().numpy() for n, v in m.state_dict().items()} for k, m in self.mods.items()}

    def _record(self, rec, tag, mod):
        """Un-clipped gradients of one optimiser step (same tags as wgg_oracle.train_batch / the product's on_step)."""
        if rec is not None:
            rec[tag] = {n: p.grad.detach().double().cpu().numpy().copy() for n, p in mod.named_parameters()}

    def train_batch(self, real, proto, noise=None, max_norm: float = 1.0, record=None, fakes=None):
        """One batch of utils.py:62-135.  ``record`` (dict) receives the un-clipped gradients of the 12 optimiser steps
        (D1_grads_i, D2_grads_i, G_grads, E_grads); ``fakes`` (dict) the two generator-side fake gestures."""
        c, tc = self.cfg, self.tc
        B = real.shape[0]
        dev = self.device
        real, proto = real.to(dev, self.dtype), proto.to(dev, self.dtype)
        it = iter(noise) if noise is not None else None
        draw = (lambda: torch.as_tensor(next(it), dtype=self.dtype, device=dev)) if it is not None else (
            lambda: torch.randn(B, c.latent_dim, dtype=self.dtype, device=dev))
        out = {}
        for it_c in range(tc.n_critic):
            for key, D in (("D1", self.D1), ("D2", self.D2)):
                with torch.no_grad():
                    z = draw() if key == "D1" else self.E(real, draw())[0]
                    fake = self.G(proto, z)
                self.opt[key].zero_grad()
                rs = D(real)
                fs = D(fake)
                loss = fs.mean() - rs.mean()
                loss.backward()
                self._record(record, f"{key}_grads_{it_c}", D)
                nn.utils.clip_grad_norm_(D.parameters(), max_norm)
                self.opt[key].step()
                out[key.lower() + "_loss"] = loss.item()
        self.opt["G"].zero_grad()
        self.opt["E"].zero_grad()
        z = draw()
        fake = self.G(proto, z)
        if fakes is not None:
            fakes["fake1"] = fake.detach().double().cpu().numpy()
        wg, ff, rf = -self.D1(fake).mean(), self.D1.feats(fake), self.D1.feats(real)
        with torch.no_grad():
            zr = self.E(fake, draw())[0]
        feat, lat = _fm(rf, ff), F.l1_loss(zr, z)
        l1 = wg + tc.lambda_feat * feat + tc.lambda_lat * lat
        out.update(cycle1_wgan=wg.item(), cycle1_feat=feat.item(), cycle1_lat=lat.item(), cycle1_total=l1.item())
        ze, mu, lv = self.E(real, draw())
        fake = self.G(proto, ze)
        if fakes is not None:
            fakes["fake2"] = fake.detach().double().cpu().numpy()
        wg, ff, rf = -self.D2(fake).mean(), self.D2.feats(fake), self.D2.feats(real)
        feat, rec = _fm(rf, ff), F.l1_loss(fake, real)
        kld = (-0.5 * torch.sum(1 + lv - mu.pow(2) - lv.exp(), dim=1)).mean()
        l2 = wg + tc.lambda_feat * feat + tc.lambda_rec * rec + tc.lambda_kld * kld
        out.update(cycle2_wgan=wg.item(), cycle2_feat=feat.item(), cycle2_rec=rec.item(), cycle2_kld=kld.item(),
                   cycle2_total=l2.item())
        (l1 + l2).backward()
        self._record(record, "G_grads", self.G)
        self._record(record, "E_grads", self.E)
        nn.utils.clip_grad_norm_(self.G.parameters(), max_norm)
        nn.utils.clip_grad_norm_(self.E.parameters(), max_norm)
        self.opt["G"].step()
        self.opt["E"].step()
        return out

    def cycle1(self, proto, real, z, eps_recover):
        """trainer.py:84-140 (train_generator_step_cycle1) with the two normal draws injected."""
        tc = self.tc
        dev = self.device
        proto, real = proto.to(dev, self.dtype), real.to(dev, self.dtype)
        z, eps_recover = torch.as_tensor(z, dtype=self.dtype, device=dev), torch.as_tensor(eps_recover, dtype=self.dtype, device=dev)
        fake = self.G(proto, z)
        wg, ff, rf = -self.D1(fake).mean(), self.D1.feats(fake), self.D1.feats(real)
        with torch.no_grad():
            zr = self.E(fake, eps_recover)[0]
        feat, lat = _fm(rf, ff), F.l1_loss(zr, z)
        total = wg + tc.lambda_feat * feat + tc.lambda_lat * lat
        return fake, total, dict(cycle1_wgan=wg.item(), cycle1_feat=feat.item(), cycle1_lat=lat.item(),
                                 cycle1_total=total.item())

    def cycle2(self, proto, real, eps):
        """trainer.py:142-193 (train_generator_step_cycle2) with the reparameterisation noise injected."""
        tc = self.tc
        dev = self.device
        proto, real = proto.to(dev, self.dtype), real.to(dev, self.dtype)
        ze, mu, lv = self.E(real, torch.as_tensor(eps, dtype=self.dtype, device=dev))
        fake = self.G(proto, ze)
        wg, ff, rf = -self.D2(fake).mean(), self.D2.feats(fake), self.D2.feats(real)
        feat, rec = _fm(rf, ff), F.l1_loss(fake, real)
        kld = (-0.5 * torch.sum(1 + lv - mu.pow(2) - lv.exp(), dim=1)).mean()
        total = wg + tc.lambda_feat * feat + tc.lambda_rec * rec + tc.lambda_kld * kld
        return fake, total, dict(cycle2_wgan=wg.item(), cycle2_feat=feat.item(), cycle2_rec=rec.item(),
                                 cycle2_kld=kld.item(), cycle2_total=total.item())

    def sample(self, proto, z):
        with torch.no_grad():
            return self.G(proto.to(self.device, self.dtype), z.to(self.device, self.dtype))
