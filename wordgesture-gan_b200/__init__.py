"""wordgesture-gan_b200: B200-native (sm_100a) implementation of the WordGesture-GAN training step.

Python host layer over libwgg_sm100.so (C ABI in include/wgg.h).  Import as ``wgg_b200`` (the directory
name contains a hyphen; the repo-root ``wgg_b200.py`` shim registers it).
"""
from ._lib import get_math_mode, set_math_mode
from .configs import DEFAULT_MODEL_CONFIG, DEFAULT_TRAINING_CONFIG, ModelConfig, TrainingConfig
from .gan_losses import (FeatureMatchingLoss, KLDivergenceLoss, LatentEncodingLoss, ReconstructionLoss,
                         WassersteinLoss, feature_matching_from_stash)
from .gan_modules import Discriminator, Generator, TemporalDiscriminator, VariationalEncoder
from .gan_trainer import WordGestureGANTrainer
from . import eval_metrics, keyboard_gpu
from .graph_step import GraphedTrainStep
from .optim import FusedClipAdam
from .resident_loader import DeviceResidentLoader
from .runner import run_training
from .train_step import log, seed_everything, train_batch, train_epoch_with_grad_clip

__all__ = [
    "ModelConfig", "TrainingConfig", "DEFAULT_MODEL_CONFIG", "DEFAULT_TRAINING_CONFIG",
    "Generator", "VariationalEncoder", "Discriminator", "TemporalDiscriminator",
    "WassersteinLoss", "FeatureMatchingLoss", "ReconstructionLoss", "LatentEncodingLoss", "KLDivergenceLoss",
    "feature_matching_from_stash", "WordGestureGANTrainer", "FusedClipAdam", "GraphedTrainStep", "DeviceResidentLoader", "run_training",
    "set_math_mode", "get_math_mode", "eval_metrics", "keyboard_gpu", "seed_everything", "log", "train_batch", "train_epoch_with_grad_clip",
]
