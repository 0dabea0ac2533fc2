"""Runs the UNMODIFIED reference's own training step / sampling call (TEST / BASELINE INFRASTRUCTURE, see
oracle/__init__.py): its WordGestureGANTrainer (src/gan/trainer.py:24-82) driven by its own
train_epoch_with_grad_clip (src/shared/utils.py:28-148) on the device asked for - "cpu" for bench.py's reference arm
and cpu_baseline, "cuda" for the library-kernel bar (cuDNN TF32 LSTM / conv + fp32 cuBLAS, torch defaults, which the
reference does not change).  The files executed are the reference's (oracle/ref_loader.py says where they come from)."""
from __future__ import annotations

import warnings

import torch

from .ref_loader import load_reference, reference_available


class ReferenceRunner:
    def __init__(self, device="cpu", seed: int = 42, batch_size: int = 512, model_kwargs=None, training_kwargs=None):
        self.ref = load_reference()
        self.device = torch.device(device)
        self.ref.utils.seed_everything(seed)
        self.model_config = self.ref.config.ModelConfig(**(model_kwargs or {}))
        self.training_config = self.ref.config.TrainingConfig(batch_size=batch_size, **(training_kwargs or {}))
        self.trainer = self.ref.trainer.WordGestureGANTrainer(self.model_config, self.training_config, device=str(self.device))

    def train_batches(self, batches, max_norm: float = 1.0):
        """One call of the reference's epoch function over ``batches`` (a list of {'gesture', 'prototype'} dicts)."""
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")  # autocast(device_type='cuda') warns on CPU-only builds
            return self.ref.utils.train_epoch_with_grad_clip(self.trainer, batches, max_norm, self.model_config,
                                                             self.training_config, self.device, scaler=None)

    def sample(self, prototype, z):
        """eval_gan.py:123-135."""
        g = self.trainer.generator
        g.eval()
        with torch.no_grad():
            out = g(prototype.to(self.device), z.to(self.device))
        g.train()
        return out


__all__ = ["ReferenceRunner", "reference_available"]
