"""Whole-batch CUDA graph of the training step.

One DataLoader batch of train_epoch_with_grad_clip is ~1500 kernel launches issued from Python (autograd
Functions -> ctypes -> C launchers); at B200 speeds the host cannot issue them as fast as the GPU retires them.
Every shape in the step is static, the noise is drawn on the device, the optimiser's step counter and learning
rate live in device memory (csrc/optim.cu: wgg_clip_adam_dev) and no call synchronises or allocates outside
torch's caching allocator - so the entire batch (12 optimiser steps, forward + backward of all four networks, the
NCCL all-reduces under data parallelism) is captured once and replayed with a single launch.
"""
from __future__ import annotations

from typing import Dict

import torch

from .train_step import train_batch

_OPTS = ("optimizer_G", "optimizer_E", "optimizer_D1", "optimizer_D2")
_MODS = ("generator", "encoder", "discriminator_1", "discriminator_2")


class GraphedTrainStep:
    def __init__(self, trainer, batch_size: int, max_norm: float, warmup: int = 2):
        self.trainer = trainer
        self.max_norm = max_norm
        self.batch_size = batch_size
        dev = trainer.device
        mc = trainer.model_config
        self.real = torch.zeros(batch_size, mc.seq_length, mc.input_dim, device=dev)
        self.proto = torch.zeros_like(self.real)
        # Warm-up (sizes the workspace, creates optimiser state, sets kernel attributes) must not disturb training
        # state: snapshot parameters / buffers / moments / counters and restore them afterwards.
        snap = self._snapshot()
        gen = torch.Generator(device=dev).manual_seed(0)
        self.real.copy_(torch.rand(self.real.shape, device=dev, generator=gen) * 2 - 1)
        self.proto.copy_(torch.rand(self.proto.shape, device=dev, generator=gen) * 2 - 1)
        from . import _lib
        # The captured launches bake scratch addresses in: the graph owns its scratch (sized by the warm-up, frozen
        # during capture, alive as long as this object) so that later eager calls cannot free it under the graph.
        self.scratch = _lib.ScratchScope()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        prev_par = getattr(trainer, "parallel_critics", "auto")
        if prev_par == "auto":
            trainer.parallel_critics = True   # warm up the two-stream critic phase the capture will use
        try:
            with torch.cuda.stream(side), self.scratch:
                for _ in range(max(1, warmup)):
                    train_batch(trainer, self.real, self.proto, max_norm)
        finally:
            trainer.parallel_critics = prev_par
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self._restore(snap)
        for name in _OPTS:
            getattr(trainer, name).refresh_device_scalars(force_step=True)
        l0 = _lib.launch_count(dev)
        self.graph = torch.cuda.CUDAGraph()
        self.scratch.frozen = True
        with self.scratch, torch.cuda.graph(self.graph):
            self.out = train_batch(trainer, self.real, self.proto, max_norm)
        self.launches_per_step = _lib.launch_count(dev) - l0  # libwgg_sm100 kernels captured (replayed every call)
        # capture records, it does not execute: training state is untouched

    def _snapshot(self):
        t = self.trainer
        snap = {"rng": torch.cuda.get_rng_state(t.device), "mods": [], "opts": []}
        for m in _MODS:
            mod = getattr(t, m)
            fb = mod.flat_buffers()
            snap["mods"].append((mod.flat_params().clone(), None if fb is None else fb.clone()))
        for o in _OPTS:
            opt = getattr(t, o)
            snap["opts"].append((opt._step, None if opt._m is None else opt._m.clone(),
                                 None if opt._v is None else opt._v.clone()))
        return snap

    def _restore(self, snap):
        t = self.trainer
        with torch.no_grad():
            for m, (p, b) in zip(_MODS, snap["mods"]):
                mod = getattr(t, m)
                mod.flat_params().copy_(p)
                if b is not None:
                    mod.flat_buffers().copy_(b)
            for o, (step, mm, vv) in zip(_OPTS, snap["opts"]):
                opt = getattr(t, o)
                opt._step = step
                if mm is not None:
                    opt._m.copy_(mm)
                    opt._v.copy_(vv)
                elif opt._m is not None:
                    opt._m.zero_()
                    opt._v.zero_()
                for st in opt.state.values():
                    st["step"].fill_(float(step))
                opt.zero_grad()
        torch.cuda.set_rng_state(snap["rng"], t.device)

    def __call__(self, real: torch.Tensor, proto: torch.Tensor) -> Dict[str, torch.Tensor]:
        """Runs one batch.  The returned tensors are the graph's static outputs (overwritten by the next call)."""
        t = self.trainer
        self.real.copy_(real, non_blocking=True)
        self.proto.copy_(proto, non_blocking=True)
        for name in _OPTS:
            getattr(t, name).refresh_device_scalars()
        self.graph.replay()
        n = t.training_config.n_critic
        t.optimizer_D1.note_graph_replays(n)
        t.optimizer_D2.note_graph_replays(n)
        t.optimizer_G.note_graph_replays(1)
        t.optimizer_E.note_graph_replays(1)
        return self.out
