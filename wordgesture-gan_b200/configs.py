"""Configuration records of the hot path.

Field names and defaults follow the reference's dataclasses (src/shared/config.py:11-33 ModelConfig,
:36-66 TrainingConfig) so that objects built by reference-side driver code (TRAIN_SCRIPT,
train_gan.py:75-78) can be passed in unchanged: every consumer in this package reads plain attributes
(duck typing), it never checks the class.
"""
from dataclasses import dataclass
from typing import Tuple


@dataclass
class ModelConfig:
    seq_length: int = 128
    input_dim: int = 3
    latent_dim: int = 32
    gen_hidden_dim: int = 48
    gen_num_layers: int = 4
    disc_hidden_dims: Tuple[int, ...] = (192, 96, 48, 24)
    use_temporal_disc: bool = True
    prototype_has_time: bool = False
    enc_hidden_dims: Tuple[int, ...] = (192, 96, 48, 32)


@dataclass
class TrainingConfig:
    batch_size: int = 512
    learning_rate: float = 2e-4
    num_epochs: int = 200
    num_workers: int = 8
    n_critic: int = 5
    lr_scheduler_eta_min: float = 1e-5
    grad_clip_norm: float = 1.0
    lambda_feat: float = 1.0
    lambda_rec: float = 4.0
    lambda_lat: float = 0.5
    lambda_kld: float = 0.02
    max_samples_per_word: int = 5
    train_ratio: float = 0.8
    save_every: int = 10
    log_every: int = 100


DEFAULT_MODEL_CONFIG = ModelConfig()
DEFAULT_TRAINING_CONFIG = TrainingConfig()
