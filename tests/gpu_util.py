"""Helpers for the -m gpu parity tests: build product modules from oracle-style state dicts."""
import numpy as np
import torch

import wgg_b200 as wgg
from golden_util import MODS
from oracle import wgg_oracle as O

ATTR = dict(G="generator", E="encoder", D1="discriminator_1", D2="discriminator_2")
DEV = "cuda:0"


def to_t(a, dev=DEV):
    return torch.from_numpy(np.ascontiguousarray(np.asarray(a, np.float32))).to(dev)


def to_np(t):
    return t.detach().double().cpu().numpy()


def model_cfg(ocfg: O.ModelCfg) -> wgg.ModelConfig:
    return wgg.ModelConfig(seq_length=ocfg.seq_length, input_dim=ocfg.input_dim, latent_dim=ocfg.latent_dim,
                           gen_hidden_dim=ocfg.gen_hidden_dim, gen_num_layers=ocfg.gen_num_layers,
                           disc_hidden_dims=tuple(ocfg.disc_hidden_dims), use_temporal_disc=ocfg.use_temporal_disc,
                           prototype_has_time=ocfg.prototype_has_time, enc_hidden_dims=tuple(ocfg.enc_hidden_dims))


def load_state(module, state):
    module.load_state_dict({k: torch.from_numpy(np.asarray(v, np.float32)) for k, v in state.items()})


def state_of(module):
    """fp32-rounded float64 copy of a module's state_dict (what the oracle should start from)."""
    return {k: v.detach().double().cpu().numpy() for k, v in module.state_dict().items()}


def trainer_from_golden(g):
    mc = wgg.ModelConfig(**g.cfg_kwargs())
    tr = wgg.WordGestureGANTrainer(mc, wgg.TrainingConfig(), DEV)
    for m in MODS:
        load_state(getattr(tr, ATTR[m]), g.init_state(m))
    return tr


def grads_of(module):
    return {k: to_np(p.grad) for k, p in module.named_parameters()}


def rand_inputs(ocfg, B, seed):
    rng = np.random.default_rng(seed)
    f32 = lambda a: a.astype(np.float32).astype(np.float64)
    real = f32(rng.uniform(-1, 1, (B, ocfg.seq_length, ocfg.input_dim)))
    proto = f32(rng.uniform(-1, 1, (B, ocfg.seq_length, ocfg.input_dim)))
    z = f32(rng.standard_normal((B, ocfg.latent_dim)))
    return real, proto, z
