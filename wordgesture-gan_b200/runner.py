"""Local training driver equal to the reference's TRAIN_SCRIPT body (train_gan.py:60-200),
without the Modal / W&B / matplotlib plumbing: same seeding, trainer, four CosineAnnealingLR schedulers
(T_max = num_epochs, eta_min = 1e-5), resume from ``latest.pt`` with scheduler replay, per-epoch
``train_epoch_with_grad_clip`` and the same checkpoint files (``latest.pt`` + ``epoch_N.pt`` in the
``get_modal_checkpoint_dict`` format, loadable by the reference and back).  Inputs are tensors (the reference's
dataset code is out of scope); they are kept on the device (DeviceResidentLoader).  Under torch.distributed the
dataset is sharded by rank, gradients are all-reduced by DataParallelGAN and only rank 0 writes checkpoints.
"""
from __future__ import annotations

import time
from pathlib import Path
from typing import Dict, List, Optional

import torch
from torch.optim.lr_scheduler import CosineAnnealingLR

from .configs import ModelConfig, TrainingConfig
from .gan_trainer import WordGestureGANTrainer
from .resident_loader import DeviceResidentLoader
from .train_step import log, seed_everything, train_epoch_with_grad_clip


def run_training(gestures: torch.Tensor, prototypes: torch.Tensor, num_epochs: int, checkpoint_dir,
                 resume: bool = True, model_config: Optional[ModelConfig] = None,
                 training_config: Optional[TrainingConfig] = None, checkpoint_every: int = 10,
                 grad_clip_norm: float = 1.0, seed: int = 42, device="cuda", use_cuda_graph: bool = True,
                 verbose: bool = True, math_mode: Optional[str] = None) -> List[Dict[str, float]]:
    """``math_mode`` ("fp32" | "tf32" | "tf32x3", see set_math_mode) is set explicitly when given; either way the
    mode the run uses is logged (every measured number in DESIGN.md is in a tensor-core mode; the process default
    is the fp32 FMA path)."""
    import torch.distributed as dist
    from . import _lib
    if math_mode is not None:
        _lib.set_math_mode(math_mode)
    model_config = model_config or ModelConfig()
    training_config = training_config or TrainingConfig(num_epochs=num_epochs, save_every=checkpoint_every)
    device = torch.device(device)
    seed_everything(seed)                                                      # train_gan.py:72
    trainer = WordGestureGANTrainer(model_config, training_config, device=device)  # :88
    trainer.use_cuda_graph = use_cuda_graph
    rank, world = 0, 1
    if dist.is_available() and dist.is_initialized():
        from .parallel import DataParallelGAN, shard_bounds
        rank, world = dist.get_rank(), dist.get_world_size()
        DataParallelGAN(trainer)
        n = (gestures.size(0) // world) * world
        lo, hi = shard_bounds(n, rank, world)
        gestures, prototypes = gestures[lo:hi], prototypes[lo:hi]
    per_rank_batch = training_config.batch_size // world if world > 1 else training_config.batch_size
    loader = DeviceResidentLoader(gestures, prototypes, per_rank_batch, shuffle=True, drop_last=world > 1, device=device,
                                  generator=torch.Generator(device=device).manual_seed(seed))
    schedulers = {k: CosineAnnealingLR(opt, T_max=num_epochs, eta_min=1e-5)   # :95-100
                  for k, opt in (("G", trainer.optimizer_G), ("E", trainer.optimizer_E),
                                 ("D1", trainer.optimizer_D1), ("D2", trainer.optimizer_D2))}
    checkpoint_dir = Path(checkpoint_dir)
    checkpoint_dir.mkdir(parents=True, exist_ok=True)
    checkpoint_path = checkpoint_dir / "latest.pt"
    start_epoch = 0
    if resume and checkpoint_path.exists():                                    # :108-121
        ckpt = torch.load(checkpoint_path, map_location=device)
        trainer.load_modal_checkpoint(ckpt)
        start_epoch = ckpt["epoch"] + 1
        for _ in range(start_epoch):
            for sched in schedulers.values():
                sched.step()
        if verbose and rank == 0:
            log(f"Resumed from epoch {start_epoch}")
    if verbose and rank == 0:
        log(f"math mode {_lib.get_math_mode()}, world size {world}, {per_rank_batch} gestures per rank per batch")
    history: List[Dict[str, float]] = []
    for epoch in range(start_epoch, num_epochs):                               # :150-199
        trainer.current_epoch = epoch
        t0 = time.time()
        losses = train_epoch_with_grad_clip(trainer, loader, grad_clip_norm, model_config, training_config, device)
        for sched in schedulers.values():
            sched.step()
        losses = dict(losses, lr=schedulers["G"].get_last_lr()[0], epoch=epoch + 1, seconds=time.time() - t0)
        history.append(losses)
        if verbose and rank == 0:
            log(f"Epoch {epoch + 1}/{num_epochs} [{losses['seconds']:.1f}s] - D1:{losses['d1_loss']:.3f} "
                f"D2:{losses['d2_loss']:.3f} C1:{losses['cycle1_total']:.3f} C2:{losses['cycle2_total']:.3f} "
                f"LR:{losses['lr']:.6f}")
        if rank == 0 and ((epoch + 1) % checkpoint_every == 0 or epoch == num_epochs - 1):
            ckpt = trainer.get_modal_checkpoint_dict()
            torch.save(ckpt, checkpoint_dir / "latest.pt")
            torch.save(ckpt, checkpoint_dir / f"epoch_{epoch + 1}.pt")
    return history
