"""Loss functions of the training step, backed by the fused CUDA reductions in csrc/loss.cu.

Interface contract = the reference's src/gan/losses.py: ``WassersteinLoss.discriminator_loss /
generator_loss`` (:26-58), ``FeatureMatchingLoss`` (:61-93), ``ReconstructionLoss`` (:96-120),
``LatentEncodingLoss`` (:123-147), ``KLDivergenceLoss`` (:150-175) - same call signatures, each returning
a 0-dim differentiable tensor that stays on the device.
"""
from __future__ import annotations

from typing import List

import torch
import torch.nn as nn

from . import _lib


def _scalar(dev) -> torch.Tensor:
    return torch.empty((), dtype=torch.float32, device=dev)


class _MeanFn(torch.autograd.Function):
    """scale * mean(x)"""

    @staticmethod
    def forward(ctx, x, scale):
        x = x.contiguous()
        c = _lib.ctx(x.device)
        out = _scalar(x.device)
        _lib.check(_lib.lib().wgg_mean(c, _lib.ptr(x), x.numel(), float(scale), 0, _lib.ptr(out), _lib.stream(x.device)), c)
        ctx.meta = (x.shape, float(scale))
        return out

    @staticmethod
    def backward(ctx, g):
        shape, scale = ctx.meta
        dx = torch.empty(shape, dtype=torch.float32, device=g.device)
        c = _lib.ctx(g.device)
        _lib.check(_lib.lib().wgg_mean_backward(c, _lib.ptr(g.contiguous()), scale, dx.numel(), _lib.ptr(dx),
                                                _lib.stream(g.device)), c)
        return dx, None


class _L1MeanFn(torch.autograd.Function):
    """mean |a - b|  (F.l1_loss).  Gradient flows to ``a`` and, if it requires grad, to ``b``."""

    @staticmethod
    def forward(ctx, a, b):
        a, b = a.contiguous(), b.contiguous()
        if a.shape != b.shape:
            raise ValueError(f"l1 loss shape mismatch {tuple(a.shape)} vs {tuple(b.shape)}")
        c = _lib.ctx(a.device)
        out = _scalar(a.device)
        _lib.check(_lib.lib().wgg_l1_mean(c, _lib.ptr(a), _lib.ptr(b), a.numel(), 1.0, 0, _lib.ptr(out),
                                          _lib.stream(a.device)), c)
        ctx.save_for_backward(a, b)
        return out

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        c = _lib.ctx(a.device)
        lib = _lib.lib()
        g = g.contiguous()
        da = db = None
        if ctx.needs_input_grad[0]:
            da = torch.empty_like(a)
            _lib.check(lib.wgg_l1_mean_backward(c, _lib.ptr(a), _lib.ptr(b), _lib.ptr(g), 1.0, a.numel(), 0,
                                                _lib.ptr(da), _lib.stream(a.device)), c)
        if ctx.needs_input_grad[1]:
            db = torch.empty_like(b)
            _lib.check(lib.wgg_l1_mean_backward(c, _lib.ptr(b), _lib.ptr(a), _lib.ptr(g), 1.0, a.numel(), 0,
                                                _lib.ptr(db), _lib.stream(a.device)), c)
        return da, db


class _KLFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mu, log_var):
        mu, log_var = mu.contiguous(), log_var.contiguous()
        c = _lib.ctx(mu.device)
        out = _scalar(mu.device)
        B, Z = mu.shape
        _lib.check(_lib.lib().wgg_kl(c, _lib.ptr(mu), _lib.ptr(log_var), B, Z, 1.0, 0, _lib.ptr(out),
                                     _lib.stream(mu.device)), c)
        ctx.save_for_backward(mu, log_var)
        return out

    @staticmethod
    def backward(ctx, g):
        mu, log_var = ctx.saved_tensors
        c = _lib.ctx(mu.device)
        B, Z = mu.shape
        dmu = torch.empty_like(mu)
        dlv = torch.empty_like(log_var)
        _lib.check(_lib.lib().wgg_kl_backward(c, _lib.ptr(mu), _lib.ptr(log_var), _lib.ptr(g.contiguous()), 1.0, B, Z,
                                              _lib.ptr(dmu), _lib.ptr(dlv), _lib.stream(mu.device)), c)
        return dmu, dlv


class _FeatureMatchingStashFn(torch.autograd.Function):
    """Feature-matching loss straight from two discriminator stashes (kernel layout, no re-layout):
    (1/K) sum_k mean|fake_k - real_k| / n_k.  The real side is treated as a constant (losses.py:91 detaches it)."""

    @staticmethod
    def forward(ctx, real_stash, fake_stash, config, B):
        c = _lib.ctx(fake_stash.device)
        cfg = _lib.c_cfg(config)
        out = _scalar(fake_stash.device)
        _lib.check(_lib.lib().wgg_feature_matching(c, cfg, _lib.ptr(real_stash), _lib.ptr(fake_stash), B, 1.0, 0,
                                                   _lib.ptr(out), _lib.stream(fake_stash.device)), c)
        ctx.meta = (config, B)
        ctx.save_for_backward(real_stash, fake_stash)
        return out

    @staticmethod
    def backward(ctx, g):
        config, B = ctx.meta
        real_stash, fake_stash = ctx.saved_tensors
        c = _lib.ctx(fake_stash.device)
        cfg = _lib.c_cfg(config)
        # only the feature blocks are written; the pooled block keeps zero gradient
        d = torch.zeros_like(fake_stash)
        _lib.check(_lib.lib().wgg_feature_matching_backward(c, cfg, _lib.ptr(real_stash), _lib.ptr(fake_stash),
                                                            _lib.ptr(g.contiguous()), 1.0, B, _lib.ptr(d),
                                                            _lib.stream(fake_stash.device)), c)
        return None, d, None, None


def feature_matching_from_stash(real_stash: torch.Tensor, fake_stash: torch.Tensor, config, batch_size: int) -> torch.Tensor:
    """Fused feature-matching loss on the raw stashes returned by ``disc.features_stash(x)``."""
    return _FeatureMatchingStashFn.apply(real_stash.detach(), fake_stash, config, batch_size)


class WassersteinLoss:
    """D: E[D(G(z))] - E[D(x)];  G: -E[D(G(z))]."""

    @staticmethod
    def discriminator_loss(real_scores: torch.Tensor, fake_scores: torch.Tensor) -> torch.Tensor:
        return _MeanFn.apply(fake_scores, 1.0) + _MeanFn.apply(real_scores, -1.0)

    @staticmethod
    def generator_loss(fake_scores: torch.Tensor) -> torch.Tensor:
        return _MeanFn.apply(fake_scores, -1.0)


class FeatureMatchingLoss(nn.Module):
    """(1/K) sum_k l1_mean(fake_k, real_k.detach()) / n_k with n_k the per-sample feature count."""

    def forward(self, real_features: List[torch.Tensor], fake_features: List[torch.Tensor]) -> torch.Tensor:
        total = None
        for real, fake in zip(real_features, fake_features):
            n_k = real.numel() / real.size(0)
            term = _L1MeanFn.apply(fake, real.detach()) / n_k
            total = term if total is None else total + term
        return total / len(real_features)


class ReconstructionLoss(nn.Module):
    def forward(self, real: torch.Tensor, fake: torch.Tensor) -> torch.Tensor:
        return _L1MeanFn.apply(fake, real)


class LatentEncodingLoss(nn.Module):
    def forward(self, z_original: torch.Tensor, z_recovered: torch.Tensor) -> torch.Tensor:
        return _L1MeanFn.apply(z_recovered, z_original)


class KLDivergenceLoss(nn.Module):
    def forward(self, mu: torch.Tensor, log_var: torch.Tensor) -> torch.Tensor:
        return _KLFn.apply(mu, log_var)
