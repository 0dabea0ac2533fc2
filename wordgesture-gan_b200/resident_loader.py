"""Device-resident input pipeline.

The reference feeds the step from ``DataLoader(num_workers=8)`` with a dict collate and one H2D copy per batch
(src/shared/data.py:526-533, src/shared/utils.py:64-65).  At > 100 k gestures/s that Python pipeline is the
bottleneck, and the whole dataset (~30 k gestures x 2 x 1.5 KB = 92 MB) fits in HBM thousands of times over.  This
loader keeps both tensors on the device, shuffles with an on-device permutation and yields the same batch dicts
(``{'gesture': (B,T,3), 'prototype': (B,T,3)}``), so ``train_epoch_with_grad_clip(trainer, loader, ...)`` runs
unchanged - with no host work per batch other than launching two gathers.
"""
from __future__ import annotations

from typing import Dict, Iterator, Optional, Sequence

import torch


class DeviceResidentLoader:
    def __init__(self, gestures: torch.Tensor, prototypes: torch.Tensor, batch_size: int, shuffle: bool = True,
                 drop_last: bool = False, device=None, generator: Optional[torch.Generator] = None,
                 words: Optional[Sequence[str]] = None):
        if gestures.shape != prototypes.shape or gestures.dim() != 3:
            raise ValueError(f"gestures {tuple(gestures.shape)} and prototypes {tuple(prototypes.shape)} must be "
                             "equal-shaped (N, T, C) tensors")
        if batch_size <= 0:
            raise ValueError("batch_size must be positive")
        if words is not None and len(words) != gestures.size(0):
            raise ValueError("words must have one entry per gesture")
        device = torch.device(device) if device is not None else gestures.device
        self.gestures = gestures.to(device=device, dtype=torch.float32).contiguous()
        self.prototypes = prototypes.to(device=device, dtype=torch.float32).contiguous()
        self.batch_size, self.shuffle, self.drop_last = int(batch_size), bool(shuffle), bool(drop_last)
        self.generator = generator
        self.words = list(words) if words is not None else None
        self.device = device

    def __len__(self) -> int:
        n = self.gestures.size(0)
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def __iter__(self) -> Iterator[Dict[str, torch.Tensor]]:
        n = self.gestures.size(0)
        perm = torch.randperm(n, device=self.device, generator=self.generator) if self.shuffle else None
        host_perm = perm.tolist() if (perm is not None and self.words is not None) else None
        for lo in range(0, n, self.batch_size):
            hi = min(lo + self.batch_size, n)
            if self.drop_last and hi - lo < self.batch_size:
                return
            if perm is None:
                batch = {"gesture": self.gestures[lo:hi], "prototype": self.prototypes[lo:hi]}
            else:
                idx = perm[lo:hi]
                batch = {"gesture": self.gestures.index_select(0, idx), "prototype": self.prototypes.index_select(0, idx)}
            if self.words is not None:
                ids = host_perm[lo:hi] if host_perm is not None else range(lo, hi)
                batch["word"] = [self.words[i] for i in ids]
            yield batch
