"""Fused gradient-clip + Adam optimiser over a FlatModule's flat buffers (csrc/optim.cu).

Replaces the pair ``torch.nn.utils.clip_grad_norm_(module.parameters(), max_norm)`` +
``torch.optim.Adam.step()`` of the reference step (src/shared/utils.py:87-88,108-109,132-135; Adam built at
src/gan/trainer.py:60-79).  It is a ``torch.optim.Optimizer``: ``param_groups[0]['lr']`` is what
``CosineAnnealingLR`` mutates (train_gan.py:95-100) and ``state_dict()`` uses torch.optim.Adam's layout
(per-parameter ``step`` / ``exp_avg`` / ``exp_avg_sq``) so reference checkpoints load both ways
(src/gan/trainer.py:208-211,226-229).
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib


class FusedClipAdam(torch.optim.Optimizer):
    def __init__(self, module, lr: float = 2e-4, betas=(0.5, 0.999), eps: float = 1e-8):
        self.module = module
        defaults = dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=0, amsgrad=False, maximize=False,
                        foreach=None, capturable=False, differentiable=False, fused=None)
        super().__init__(list(module.parameters()), defaults)
        self._m = None
        self._v = None
        self._step = 0
        self._step_dev = None   # int32 device counter (what the kernels read; CUDA-graph capturable)
        self._lr_dev = None     # float32 device scalar mirroring param_groups[0]['lr']
        self._lr_cached = None
        self.process_group = None  # set by parallel.DataParallelGAN: all-reduce (mean) grads before clipping
        self.world_size = 1
        self.p2p = None            # parallel.P2PBucket: one-shot peer-memory reduce instead of the NCCL collective
        self.last_grad_norm = None  # device scalar: pre-clip global L2 norm of the last step

    # ---- flat views ---------------------------------------------------------------------------
    def _params(self):
        return self.param_groups[0]["params"]

    def _ensure_state(self):
        flat = self.module.flat_params()
        if self._m is None or self._m.device != flat.device or self._m.numel() != flat.numel():
            old_m, old_v = self._m, self._v
            self._m = torch.zeros_like(flat)
            self._v = torch.zeros_like(flat)
            if old_m is not None and old_m.numel() == flat.numel():
                self._m.copy_(old_m)
                self._v.copy_(old_v)
        if len(self.state) != len(self._params()) or not self._state_is_aliased():
            self._alias_state()
        if self._step_dev is None or self._step_dev.device != flat.device:
            self._step_dev = torch.full((1,), self._step, dtype=torch.int32, device=flat.device)
            self._lr_dev = torch.zeros((1,), dtype=torch.float32, device=flat.device)
            self._lr_cached = None
        return flat

    def refresh_device_scalars(self, force_step: bool = False):
        """Mirror the host-side learning rate (mutated by LR schedulers) and, on request, the step count onto the
        device scalars the kernels read.  Called before every eager step and before every graph replay."""
        lr = float(self.param_groups[0]["lr"])
        if lr != self._lr_cached:
            self._lr_dev.fill_(lr)
            self._lr_cached = lr
        if force_step:
            self._step_dev.fill_(self._step)

    def note_graph_replays(self, n_steps: int):
        """A captured graph advanced the device step counter by ``n_steps``: keep the host bookkeeping in sync."""
        self._step += n_steps
        for st in self.state.values():
            st["step"].fill_(float(self._step))

    def _state_is_aliased(self) -> bool:
        off = self._m.data_ptr()
        for p in self._params():
            st = self.state.get(p)
            if not st or st["exp_avg"].data_ptr() != off:
                return False
            off += 4 * p.numel()
        return True

    def _alias_state(self):
        """(Re)publish the flat moments as torch.optim.Adam-shaped per-parameter state."""
        off = 0
        for p in self._params():
            n = p.numel()
            st = self.state.get(p)
            if st and "exp_avg" in st and st["exp_avg"].data_ptr() != self._m.data_ptr() + 4 * off:
                # state arrived from load_state_dict: import it
                self._m[off:off + n].copy_(st["exp_avg"].reshape(-1))
                self._v[off:off + n].copy_(st["exp_avg_sq"].reshape(-1))
                self._step = int(st["step"]) if "step" in st else self._step
            self.state[p] = {
                "step": torch.tensor(float(self._step)),
                "exp_avg": self._m[off:off + n].view(p.shape),
                "exp_avg_sq": self._v[off:off + n].view(p.shape),
            }
            off += n

    def zero_grad(self, set_to_none: bool = True) -> None:
        """One fill of the module's flat gradient buffer (FlatModule.zero_grad) instead of dropping every .grad."""
        self.module.zero_grad(set_to_none)

    def flat_grad(self) -> torch.Tensor:
        """Gradients as ONE contiguous tensor in parameter order.  The backward kernels already write one flat
        buffer per module (each .grad is a view of it); otherwise the pieces are packed here and .grad re-aliased."""
        params = self._params()
        grads = [p.grad for p in params]
        total = sum(p.numel() for p in params)
        g0 = grads[0]
        if g0 is not None and g0.is_contiguous():
            ok = True
            base = g0.data_ptr()
            off = 0
            for p, g in zip(params, grads):
                if g is None or g.dtype != torch.float32 or not g.is_contiguous() or g.data_ptr() != base + 4 * off:
                    ok = False
                    break
                off += p.numel()
            if ok and g0.untyped_storage().nbytes() - 4 * g0.storage_offset() >= 4 * total:
                return g0.as_strided((total,), (1,), g0.storage_offset())
        flat = torch.zeros(total, dtype=torch.float32, device=params[0].device)
        off = 0
        for p, g in zip(params, grads):
            n = p.numel()
            if g is not None:
                flat[off:off + n].copy_(g.reshape(-1))
            p.grad = flat[off:off + n].view(p.shape)
            off += n
        return flat

    # ---- step ---------------------------------------------------------------------------------
    @torch.no_grad()
    def step(self, closure=None, max_norm: Optional[float] = None, grads_already_reduced: bool = False):
        """One fused update.  ``max_norm`` > 0 applies clip_grad_norm_ semantics (L2, coefficient
        min(1, max_norm / (norm + 1e-6))) inside the same pass.  Under data parallelism the flat gradient bucket is
        mean-all-reduced first, unless the caller already did (``grads_already_reduced``: train_batch reduces the
        generator's and the encoder's buckets with ONE collective)."""
        loss = closure() if closure is not None else None
        group = self.param_groups[0]
        flat = self._ensure_state()
        g = self.flat_grad()
        if self.process_group is not None and self.world_size > 1 and not grads_already_reduced:
            if self.p2p is not None and g.data_ptr() == self.p2p.shared.data_ptr():
                self.p2p.allreduce_mean_()
            else:
                from .parallel import allreduce_mean_
                allreduce_mean_(g, self.process_group, self.world_size)
        dev = flat.device
        c = _lib.ctx(dev)
        lib = _lib.lib()
        ws = _lib.workspace(dev, lib.wgg_clip_adam_workspace_floats())
        if self.last_grad_norm is None or self.last_grad_norm.device != dev:
            self.last_grad_norm = torch.zeros((), dtype=torch.float32, device=dev)
        capturing = torch.cuda.is_current_stream_capturing()
        if not capturing:
            self.refresh_device_scalars()
        b1, b2 = group["betas"]
        _lib.check(lib.wgg_clip_adam_dev(c, _lib.ptr(flat), _lib.ptr(g), _lib.ptr(self._m), _lib.ptr(self._v), flat.numel(),
                                         self._lr_dev.data_ptr(), float(b1), float(b2), float(group["eps"]),
                                         self._step_dev.data_ptr(), float(max_norm) if max_norm else 0.0,
                                         _lib.ptr(self.last_grad_norm), _lib.ptr(ws), _lib.stream(dev)), c)
        if capturing:
            return loss  # host bookkeeping is done per replay (note_graph_replays)
        self._step += 1
        for st in self.state.values():
            st["step"].fill_(float(self._step))
        return loss

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        if self.state:
            any_state = next(iter(self.state.values()))
            self._step = int(any_state["step"]) if "step" in any_state else 0
            flat = self.module.flat_params()
            self._m = torch.zeros_like(flat)
            self._v = torch.zeros_like(flat)
            self._alias_state()
            if self._step_dev is not None:
                self._step_dev.fill_(self._step)
