"""The training step: ``train_epoch_with_grad_clip`` and the seed / log helpers.

Interface contract = src/shared/utils.py:12-148 of the reference.  Per DataLoader batch:
n_critic x { D1 step on G(proto, randn) ; D2 step on G(proto, E(real)) } then one joint G/E step on
cycle-1 + cycle-2 losses; every optimiser step is clip_grad_norm_(max_norm) + Adam, here one fused kernel.
Loss scalars are accumulated on the device and read back once per epoch (the reference syncs 11 times per
batch; its return contract is only the epoch means, utils.py:138-148).
"""
from __future__ import annotations

import random
from typing import Dict, List, Optional

import numpy as np
import torch

from . import _lib
from .gan_losses import WassersteinLoss


def seed_everything(seed: int) -> None:
    """python, numpy, torch CPU and current-CUDA-device generators (utils.py:12-20)."""
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed(seed)


def log(msg: str) -> None:
    print(msg, flush=True)


# A/B switch for the two-call critic-phase generator (see train_batch); on unless WGG_SPLIT_GEN=0
_SPLIT_DEFAULT = __import__("os").environ.get("WGG_SPLIT_GEN", "1") != "0"


def n_noise_draws(training_config) -> int:
    return 2 * training_config.n_critic + 3


def _critic_side_stream(trainer, dev):
    st = getattr(trainer, "_critic_stream", None)
    if st is None:
        st = trainer._critic_stream = torch.cuda.Stream(device=dev)
    return st


def dp_slice(trainer):
    """(rank, world) of the data-parallel group attached to ``trainer`` by parallel.DataParallelGAN, else (0, 1)."""
    dp = getattr(trainer, "_dp", None)
    return (dp.rank, dp.world_size) if dp is not None else (0, 1)


def train_batch(trainer, real_gesture: torch.Tensor, prototype: torch.Tensor, max_norm: float,
                noise: Optional[List[torch.Tensor]] = None, on_step=None, training_config=None,
                model_config=None) -> Dict[str, torch.Tensor]:
    """One batch of the step.  ``noise`` optionally injects the 2*n_critic+3 (B, Z) normal draws in consumption
    order (z_rand, eps) x n_critic, z, eps_recover, eps; by default each is a torch.randn call made at the point
    where the reference makes it, so the global RNG stream is consumed identically.
    Under data parallelism (a DataParallelGAN is attached) ``real_gesture`` / ``prototype`` are this rank's shard and
    every default draw is made for the GLOBAL batch - torch.randn(world * B, Z) on every rank from the shared seed -
    and sliced to the rank's rows, so an N-rank run consumes the same numbers as the 1-rank run on the global batch
    (injected ``noise`` is taken as already sharded).
    ``training_config`` / ``model_config`` default to the trainer's (train_epoch_with_grad_clip forwards its own
    arguments, utils.py:68,71).  ``on_step(tag, optimizer)`` (optional) is called right before each of the 12
    optimiser steps (tests use it to read the un-clipped gradients).  Returns 0-dim device tensors for all 11
    logged scalars."""
    tc = training_config if training_config is not None else trainer.training_config
    mc = model_config if model_config is not None else trainer.model_config
    dev = real_gesture.device
    B = real_gesture.size(0)
    it = iter(noise) if noise is not None else None
    rank, world = dp_slice(trainer)

    def draw():
        if it is not None:
            return next(it)
        from .parallel import randn_rank_rows
        return randn_rank_rows(B, mc.latent_dim, rank, world, dev)

    out: Dict[str, torch.Tensor] = {}
    # ---- critic phase (utils.py:68-109) ----
    # Neither the generator nor the encoder is updated inside the critic loop, so all 2*n_critic fake batches depend
    # only on (prototype, real, noise): draw the noise in the reference's order (z_rand_i then eps_i per iteration)
    # and run ONE generator call on the stacked batch and ONE encoder call - identical numbers, far better SM fill
    # for the persistent recurrent kernel, a fifth of the launches.
    n = tc.n_critic
    fakes_1 = fakes_2 = None
    # The D1 chain and the D2 chain are independent (different networks, optimisers and fake batches; they only read
    # `real`), so D2's steps are issued on a second stream: its many small launches (Linear layers, spectral norm,
    # reductions) overlap D1's bandwidth-bound conv kernels and vice versa.  Inside a captured CUDA graph the two
    # chains become parallel branches.  Results are identical to the sequential order.
    # Eager execution is bound by host launch issue, where a second stream only adds synchronisation calls, so the
    # side stream is used under graph capture (and by the capture's warm-up) only.
    main = torch.cuda.current_stream(dev) if dev.type == "cuda" else None
    par = getattr(trainer, "parallel_critics", "auto")
    use_side = main is not None and n > 0 and (par is True or (par == "auto" and torch.cuda.is_current_stream_capturing()))
    side = _critic_side_stream(trainer, dev) if use_side else None
    capturing = main is not None and torch.cuda.is_current_stream_capturing()
    if n > 0:
        zs, epss = [], []
        for _ in range(n):
            zs.append(draw())
            epss.append(draw())
        with torch.no_grad():
            if n > 1:
                # the encoder is deterministic up to the reparameterisation (models.py:80-86): mu / log_var of `real`
                # are the same in every critic iteration, only eps differs - one encoder pass, n re-parameterisations
                z0, mu, log_var = trainer.encoder(real_gesture, epss[0])
                std = torch.exp(0.5 * log_var)
                z_enc = torch.cat([z0] + [torch.addcmul(mu, e, std) for e in epss[1:]], 0)
            else:
                z_enc, _, _ = trainer.encoder(real_gesture, epss[0])
            z_rand = torch.cat(zs, 0) if n > 1 else zs[0]
            proto_rep = prototype.repeat(n, 1, 1) if n > 1 else prototype
            if side is not None and getattr(trainer, "split_critic_generator", _SPLIT_DEFAULT):
                # two generator calls of n*B gestures, one per critic chain, each on its chain's stream: a layer launch
                # of the persistent recurrent kernel is (n*B/128) x 2 CTAs, rarely a whole number of waves of the 148
                # SMs - with two independent calls in flight the tail wave of one call's layer is filled by the other
                # call's CTAs (10 x 4096 gestures: 20 waves -> 17.3), and D1's chain starts as soon as ITS fakes exist
                side.wait_stream(main)
                fake_1 = trainer.generator(proto_rep, z_rand)
                with torch.cuda.stream(side), _lib.lane(dev, 1):
                    fake_2 = trainer.generator(proto_rep, z_enc)
                if not capturing:
                    for t_ in (proto_rep, z_enc):
                        t_.record_stream(side)
            else:
                fake_all = trainer.generator(torch.cat([proto_rep, proto_rep], 0), torch.cat([z_rand, z_enc], 0))
                fake_1, fake_2 = fake_all[:n * B], fake_all[n * B:]
                if side is not None:
                    side.wait_stream(main)
        fakes_1 = fake_1.view(n, B, *fake_1.shape[1:])
        fakes_2 = fake_2.view(n, B, *fake_2.shape[1:])

    def d_step(name, disc, opt, fake, critic_it):
        opt.zero_grad()
        real_scores = disc(real_gesture)
        fake_scores = disc(fake)
        loss = WassersteinLoss.discriminator_loss(real_scores, fake_scores)
        loss.backward()
        if on_step is not None:
            on_step(f"{name[:2].upper()}_grads_{critic_it}", opt)
        opt.step(max_norm=max_norm)
        out[name] = loss.detach()
        if side is not None and name == "d2_loss" and not torch.cuda.is_current_stream_capturing():
            # eager two-stream mode: this scalar was allocated on the side stream and is read on the main one
            out[name].record_stream(main)

    for critic_it in range(n):
        d_step("d1_loss", trainer.discriminator_1, trainer.optimizer_D1, fakes_1[critic_it], critic_it)
        if side is not None:
            with torch.cuda.stream(side), _lib.lane(dev, 1):
                d_step("d2_loss", trainer.discriminator_2, trainer.optimizer_D2, fakes_2[critic_it], critic_it)
        else:
            d_step("d2_loss", trainer.discriminator_2, trainer.optimizer_D2, fakes_2[critic_it], critic_it)
    if side is not None:
        main.wait_stream(side)

    trainer.optimizer_G.zero_grad()
    trainer.optimizer_E.zero_grad()
    z_c1, eps_rec, eps_c2 = draw(), draw(), draw()  # reference order: cycle-1 z, its recovery eps, cycle-2 eps
    _, _, loss1, loss2, d1, d2 = trainer.cycles_tensors(prototype, real_gesture, z=z_c1, eps_recover=eps_rec, eps=eps_c2,
                                                        training_config=tc)
    # the discriminators' own weight gradients of this backward are never used (the reference zeroes them before
    # the next critic step, utils.py:75,96): skip computing them, keep d(loss)/d(fake gesture)
    _lib.SKIP_DISC_WEIGHT_GRADS = True
    try:
        (loss1 + loss2).backward()
    finally:
        _lib.SKIP_DISC_WEIGHT_GRADS = False
    if on_step is not None:
        on_step("G_grads", trainer.optimizer_G)
        on_step("E_grads", trainer.optimizer_E)
    reduced = False
    if world > 1 and trainer.optimizer_G.process_group is not None and trainer.optimizer_G.p2p is None:
        # NCCL path: one collective for the generator's and the encoder's gradient buckets (G || E = 301 523 floats,
        # 1.2 MB); the peer-memory path reduces each bucket with its own one-shot launch inside step()
        from .parallel import allreduce_mean_
        g_g, g_e = trainer.optimizer_G.flat_grad(), trainer.optimizer_E.flat_grad()
        joint = torch.cat([g_g, g_e])
        allreduce_mean_(joint, trainer.optimizer_G.process_group, world)
        g_g.copy_(joint[:g_g.numel()])
        g_e.copy_(joint[g_g.numel():])
        reduced = True
    trainer.optimizer_G.step(max_norm=max_norm, grads_already_reduced=reduced)
    trainer.optimizer_E.step(max_norm=max_norm, grads_already_reduced=reduced)
    for d in (d1, d2):
        for k, v in d.items():
            out[k] = v.detach()
    return out


def train_epoch_with_grad_clip(trainer, dataloader, max_norm, model_config, training_config, device, scaler=None):
    """Drop-in for src/shared/utils.py:28.  ``scaler`` must be None: the reference hard-wires it to None
    (train_gan.py:90-92) and the fp32 path is the only one implemented."""
    if scaler is not None:
        raise NotImplementedError("mixed precision is a dead branch in the reference (scaler is always None)")
    for m in (trainer.generator, trainer.encoder, trainer.discriminator_1, trainer.discriminator_2):
        m.train()
    keys = ("d1_loss", "d2_loss", "cycle1_total", "cycle2_total")
    sums = None
    num_batches = 0
    use_graph = getattr(trainer, "use_cuda_graph", False)
    if use_graph == "auto":
        try:
            use_graph = torch.device(device).type == "cuda" and len(dataloader) >= 8
        except TypeError:  # an iterable without a length
            use_graph = False
    use_graph = bool(use_graph)
    # the reference reads n_critic / the loss weights / latent_dim from the ARGUMENTS (utils.py:68,71); a captured
    # graph is specific to the configuration it was captured with
    same_cfg = training_config is trainer.training_config or training_config == trainer.training_config
    for batch in dataloader:
        real = batch["gesture"].to(device, non_blocking=True)
        proto = batch["prototype"].to(device, non_blocking=True)
        if use_graph:
            # full batches replay one captured CUDA graph (graph_step.py); a ragged last batch runs eagerly
            from .graph_step import GraphedTrainStep
            cache = trainer.__dict__.setdefault("_graphed_steps", {})
            key = (real.size(0), float(max_norm), _lib.get_math_mode())
            if same_cfg and key not in cache and (not cache or real.size(0) == training_config.batch_size):
                cache[key] = GraphedTrainStep(trainer, real.size(0), max_norm)
            out = cache[key](real, proto) if (same_cfg and key in cache) else train_batch(
                trainer, real, proto, max_norm, training_config=training_config, model_config=model_config)
        else:
            out = train_batch(trainer, real, proto, max_norm, training_config=training_config, model_config=model_config)
        vals = torch.stack([out[k] for k in keys])
        sums = vals if sums is None else sums + vals
        num_batches += 1
    if num_batches == 0:
        raise ZeroDivisionError("empty dataloader")
    means = (sums / num_batches).tolist()  # the only host sync of the epoch
    # the persistent tcgen05 kernels end on a bounded mbarrier wait instead of hanging; a wedged pipeline leaves a
    # code in the context's error word - surface it here, at the epoch's one synchronisation point
    code = _lib.async_error(device)
    if code:
        raise _lib.WggError(f"a persistent tensor-core kernel timed out during this epoch (pipeline code {code}); "
                            "the epoch's updates are not trustworthy")
    if not all(np.isfinite(m) for m in means):
        raise _lib.WggError(f"non-finite epoch losses {dict(zip(keys, means))}")
    return dict(zip(keys, means))
