"""WordGestureGANTrainer drop-in: owns the four networks, the loss objects and the four optimisers.

Interface contract = src/gan/trainer.py:24-230 of the reference: same constructor, attribute names
(.generator .encoder .discriminator_1 .discriminator_2 .optimizer_G/E/D1/D2 .current_epoch ...), the two
generator-side cycle builders returning ``(fake_gesture, total_loss, loss_dict)`` and the checkpoint dict
format of get_modal_checkpoint_dict / load_modal_checkpoint.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from .configs import DEFAULT_MODEL_CONFIG, DEFAULT_TRAINING_CONFIG, ModelConfig, TrainingConfig
from .gan_losses import (FeatureMatchingLoss, KLDivergenceLoss, LatentEncodingLoss, ReconstructionLoss,
                         WassersteinLoss, feature_matching_from_stash)
from .gan_modules import Discriminator, Generator, TemporalDiscriminator, VariationalEncoder
from .optim import FusedClipAdam

_MODULES = (("generator", "optimizer_G"), ("discriminator_1", "optimizer_D1"),
            ("discriminator_2", "optimizer_D2"), ("encoder", "optimizer_E"))


class WordGestureGANTrainer:
    def __init__(self, model_config: ModelConfig = DEFAULT_MODEL_CONFIG,
                 training_config: TrainingConfig = DEFAULT_TRAINING_CONFIG,
                 device: str = "cuda" if torch.cuda.is_available() else "cpu"):
        self.model_config = model_config
        self.training_config = training_config
        self.device = torch.device(device)
        # construction order fixes the RNG stream of the initial weights: G, E, D1, D2 (trainer.py:44-51)
        self.generator = Generator(model_config).to(self.device)
        self.encoder = VariationalEncoder(model_config).to(self.device)
        disc_cls = TemporalDiscriminator if model_config.use_temporal_disc else Discriminator
        self.discriminator_1 = disc_cls(model_config).to(self.device)
        self.discriminator_2 = disc_cls(model_config).to(self.device)

        self.feature_matching_loss = FeatureMatchingLoss()
        self.reconstruction_loss = ReconstructionLoss()
        self.latent_encoding_loss = LatentEncodingLoss()
        self.kl_divergence_loss = KLDivergenceLoss()

        lr = training_config.learning_rate
        self.optimizer_G = FusedClipAdam(self.generator, lr=lr, betas=(0.5, 0.999))
        self.optimizer_E = FusedClipAdam(self.encoder, lr=lr, betas=(0.5, 0.999))
        self.optimizer_D1 = FusedClipAdam(self.discriminator_1, lr=lr, betas=(0.5, 0.999))
        self.optimizer_D2 = FusedClipAdam(self.discriminator_2, lr=lr, betas=(0.5, 0.999))
        self.current_epoch = 0
        # train_epoch_with_grad_clip replays one captured CUDA graph per full batch (graph_step.py): True / False force
        # it on / off; "auto" uses it for epochs long enough to amortise warm-up + capture (>= 8 batches)
        self.use_cuda_graph = "auto"

    # ---- generator-side cycles ------------------------------------------------------------------
    def _randn(self, B: int) -> torch.Tensor:
        """One (B, Z) normal draw; under data parallelism the global draw sliced to this rank (parallel.py)."""
        from .parallel import randn_rank_rows
        from .train_step import dp_slice
        rank, world = dp_slice(self)
        return randn_rank_rows(B, self.model_config.latent_dim, rank, world, self.device)

    def _adversarial_terms(self, disc, fake, real):
        """-mean D(fake) and the feature-matching term; three discriminator calls in the reference's order
        (score(fake), features(fake), features(real)) because each advances the spectral-norm power iteration
        (trainer.py:111-113 / :167-169)."""
        fake_scores = disc(fake)
        fake_stash = disc.features_stash(fake)
        real_stash = disc.features_stash(real)
        wgan = WassersteinLoss.generator_loss(fake_scores)
        feat = feature_matching_from_stash(real_stash, fake_stash, self.model_config, fake.shape[0])
        return wgan, feat

    def _encode_with_kl(self, real_gesture, eps):
        """Encoder pass of cycle 2 (trainer.py:161-162) together with its KL term (:171): the stock pair
        (VariationalEncoder, KLDivergenceLoss) runs as the fused latent-head kernel; a user-replaced encoder or KL
        module is called the reference's way."""
        if isinstance(self.kl_divergence_loss, KLDivergenceLoss) and hasattr(self.encoder, "forward_with_kl"):
            return self.encoder.forward_with_kl(real_gesture, eps)
        z_enc, mu, log_var = self.encoder(real_gesture, eps)
        return z_enc, mu, log_var, self.kl_divergence_loss(mu, log_var)

    def cycle1_tensors(self, prototype, real_gesture, z: Optional[torch.Tensor] = None,
                       eps_recover: Optional[torch.Tensor] = None):
        """Cycle 1 (z -> X' -> z').  Returns (fake, total, dict of 0-dim device tensors) without host syncs.
        ``z`` / ``eps_recover`` inject the two normal draws (default: torch.randn, in the reference's order)."""
        tc = self.training_config
        B = prototype.size(0)
        if z is None:
            z = self._randn(B)
        fake = self.generator(prototype, z)
        wgan, feat = self._adversarial_terms(self.discriminator_1, fake, real_gesture)
        if eps_recover is None:
            eps_recover = self._randn(B)  # drawn where the reference draws it: inside the recovery pass (trainer.py:118)
        with torch.no_grad():  # latent recovery carries no gradient in the reference either (trainer.py:116-119)
            z_rec, _, _ = self.encoder(fake, eps_recover)
        lat = self.latent_encoding_loss(z, z_rec)
        total = wgan + tc.lambda_feat * feat + tc.lambda_lat * lat
        return fake, total, {"cycle1_wgan": wgan, "cycle1_feat": feat, "cycle1_lat": lat, "cycle1_total": total}

    def cycle2_tensors(self, prototype, real_gesture, eps: Optional[torch.Tensor] = None):
        """Cycle 2 (X -> z -> X')."""
        tc = self.training_config
        if eps is None:
            eps = self._randn(prototype.size(0))
        z_enc, mu, log_var, kld = self._encode_with_kl(real_gesture, eps)
        fake = self.generator(prototype, z_enc)
        wgan, feat = self._adversarial_terms(self.discriminator_2, fake, real_gesture)
        rec = self.reconstruction_loss(real_gesture, fake)
        total = wgan + tc.lambda_feat * feat + tc.lambda_rec * rec + tc.lambda_kld * kld
        return fake, total, {"cycle2_wgan": wgan, "cycle2_feat": feat, "cycle2_rec": rec, "cycle2_kld": kld,
                             "cycle2_total": total}

    def cycles_tensors(self, prototype, real_gesture, z: Optional[torch.Tensor] = None,
                       eps_recover: Optional[torch.Tensor] = None, eps: Optional[torch.Tensor] = None,
                       training_config: Optional[TrainingConfig] = None):
        """Cycle 1 and cycle 2 together, with ONE generator call on the stacked batch [z ; z_enc] (the two cycles'
        generator passes are independent, so stacking them is exact and doubles the SM fill of the persistent
        recurrent kernels; autograd then runs one generator backward for both).  The three normal draws are made
        in the reference's order (trainer.py:105, :118 via models.py:85, :161) before any of them is consumed.
        Returns (fake1, fake2, total1, total2, dict1, dict2) with the same values as cycle1_tensors / cycle2_tensors."""
        tc = training_config if training_config is not None else self.training_config
        B = prototype.size(0)
        if z is None:
            z = self._randn(B)
        if eps_recover is None:
            eps_recover = self._randn(B)
        if eps is None:
            eps = self._randn(B)
        z_enc, mu, log_var, kld = self._encode_with_kl(real_gesture, eps)
        fake = self.generator(torch.cat([prototype, prototype], 0), torch.cat([z, z_enc], 0))
        fake1, fake2 = fake[:B], fake[B:]
        wgan1, feat1 = self._adversarial_terms(self.discriminator_1, fake1, real_gesture)
        with torch.no_grad():
            z_rec, _, _ = self.encoder(fake1, eps_recover)
        lat = self.latent_encoding_loss(z, z_rec)
        total1 = wgan1 + tc.lambda_feat * feat1 + tc.lambda_lat * lat
        wgan2, feat2 = self._adversarial_terms(self.discriminator_2, fake2, real_gesture)
        rec = self.reconstruction_loss(real_gesture, fake2)
        total2 = wgan2 + tc.lambda_feat * feat2 + tc.lambda_rec * rec + tc.lambda_kld * kld
        d1 = {"cycle1_wgan": wgan1, "cycle1_feat": feat1, "cycle1_lat": lat, "cycle1_total": total1}
        d2 = {"cycle2_wgan": wgan2, "cycle2_feat": feat2, "cycle2_rec": rec, "cycle2_kld": kld, "cycle2_total": total2}
        return fake1, fake2, total1, total2, d1, d2

    def train_generator_step_cycle1(self, prototype: torch.Tensor, real_gesture: torch.Tensor
                                    ) -> Tuple[torch.Tensor, torch.Tensor, Dict[str, float]]:
        fake, total, d = self.cycle1_tensors(prototype, real_gesture)
        return fake, total, {k: v.item() for k, v in d.items()}

    def train_generator_step_cycle2(self, prototype: torch.Tensor, real_gesture: torch.Tensor
                                    ) -> Tuple[torch.Tensor, torch.Tensor, Dict[str, float]]:
        fake, total, d = self.cycle2_tensors(prototype, real_gesture)
        return fake, total, {k: v.item() for k, v in d.items()}

    # ---- checkpoints ------------------------------------------------------------------------------
    def get_modal_checkpoint_dict(self) -> dict:
        ckpt = {"epoch": self.current_epoch}
        for mod, opt in _MODULES:
            ckpt[mod] = getattr(self, mod).state_dict()
        for mod, opt in _MODULES:
            ckpt[opt] = getattr(self, opt).state_dict()
        return ckpt

    def load_modal_checkpoint(self, checkpoint: dict):
        self.current_epoch = checkpoint["epoch"] + 1
        for mod, opt in _MODULES:
            getattr(self, mod).load_state_dict(checkpoint[mod])
        for mod, opt in _MODULES:
            getattr(self, opt).load_state_dict(checkpoint[opt])
        print(f"Loaded modal checkpoint from epoch {checkpoint['epoch'] + 1}")
