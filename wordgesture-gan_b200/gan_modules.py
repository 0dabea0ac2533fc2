"""The four networks of the training step as torch.nn.Module drop-ins backed by libwgg_sm100.so.

Interface contract = the reference's src/gan/models.py: constructor
``Cls(config)``, ``Generator.forward(prototype, z)``, ``VariationalEncoder.forward(x) -> (z, mu, log_var)``
and ``.reparameterize``, ``[Temporal]Discriminator.forward(x)`` / ``.get_all_features(x)``, identical
``state_dict`` keys/shapes/order and identical ``named_parameters`` order (Adam state indices).
Initialisation consumes the torch RNG in the same order with the same distributions, so
``seed_everything(s)`` followed by construction yields the reference's initial weights bit for bit.

Each forward is a torch.autograd.Function that calls hand-written CUDA through the C ABI; all
parameters of a module live in ONE flat fp32 buffer (the named nn.Parameters are views into it), which is
what the kernels, the fused clip+Adam step and the data-parallel all-reduce operate on.
"""
from __future__ import annotations

import math
from typing import List, Tuple

import torch
import torch.nn as nn

from . import _lib
from .configs import DEFAULT_MODEL_CONFIG, ModelConfig


# ------------------------------------------------------------------------------------------------
# flat parameter storage
# ------------------------------------------------------------------------------------------------
class FlatModule(nn.Module):
    """nn.Module whose parameters (and, separately, buffers) alias one contiguous fp32 buffer each."""

    def __init__(self):
        super().__init__()
        self._flat = None
        self._flat_buf = None
        self._gflat = None  # persistent flat gradient buffer; every parameter's .grad is a view into it

    def _flatten(self) -> None:
        with torch.no_grad():
            params = list(self.parameters())
            if params:
                dev = params[0].device
                flat = torch.empty(sum(p.numel() for p in params), dtype=torch.float32, device=dev)
                off = 0
                for p in params:
                    n = p.numel()
                    flat[off:off + n].copy_(p.detach().reshape(-1))
                    p.data = flat[off:off + n].view(p.shape)
                    p.grad = None
                    off += n
                self._flat = flat
            bufs = list(self.buffers())
            if bufs:
                fb = torch.empty(sum(b.numel() for b in bufs), dtype=torch.float32, device=bufs[0].device)
                off = 0
                for b in bufs:
                    n = b.numel()
                    fb[off:off + n].copy_(b.reshape(-1))
                    b.data = fb[off:off + n].view(b.shape)
                    off += n
                self._flat_buf = fb

    def _is_flat(self) -> bool:
        if self._flat is None:
            return False
        off = self._flat.data_ptr()
        for p in self.parameters():
            if p.data_ptr() != off or p.dtype != torch.float32:
                return False
            off += 4 * p.numel()
        if self._flat_buf is not None:
            off = self._flat_buf.data_ptr()
            for b in self.buffers():
                if b.data_ptr() != off:
                    return False
                off += 4 * b.numel()
        return True

    def flat_params(self) -> torch.Tensor:
        """The module's parameters as one contiguous tensor (named_parameters order)."""
        if not self._is_flat():
            self._flatten()
        return self._flat

    def flat_buffers(self) -> torch.Tensor:
        if not self._is_flat():
            self._flatten()
        return self._flat_buf

    # ---- gradients -------------------------------------------------------------------------------
    # The backward kernels ACCUMULATE (+=) into one flat gradient buffer per module (include/wgg.h), and that buffer
    # is what clip + Adam and the data-parallel all-reduce consume.  Handing autograd a fresh zero-filled tensor per
    # backward call would make it add the pieces into .grad with one elementwise launch per parameter (130 adds + 83
    # fills per training step); instead the module owns ONE persistent buffer, every parameter's .grad is a view of it,
    # the autograd Functions accumulate straight into it and return None for the parameter inputs.  zero_grad() is one
    # fill.  (Consequence: torch.autograd.grad(..., parameters) does not see these gradients - use .backward().)
    def grad_buffer(self) -> torch.Tensor:
        """The flat gradient buffer, with every parameter's .grad attached as a view of it.  A parameter whose .grad is
        None (never written, or reset by zero_grad(set_to_none=True)) gets its slice zeroed; a foreign .grad tensor
        (assigned by the user) is copied in - so accumulation semantics are those of autograd."""
        flat = self.flat_params()
        if self._gflat is None or self._gflat.device != flat.device or self._gflat.numel() != flat.numel():
            self._gflat = torch.zeros_like(flat)
        off = 0
        base = self._gflat.data_ptr()
        with torch.no_grad():
            for p in self.parameters():
                n = p.numel()
                g = p.grad
                if g is None:
                    self._gflat[off:off + n].zero_()
                    p.grad = self._gflat[off:off + n].view(p.shape)
                elif g.data_ptr() != base + 4 * off or g.dtype != torch.float32:
                    self._gflat[off:off + n].copy_(g.reshape(-1))
                    p.grad = self._gflat[off:off + n].view(p.shape)
                off += n
        return self._gflat

    def zero_grad(self, set_to_none: bool = True) -> None:
        """One fill of the flat gradient buffer; the .grad views stay attached (set_to_none is accepted for signature
        compatibility: a zeroed gradient and an absent one are the same to every consumer in this package)."""
        if self._gflat is None:
            return super().zero_grad(set_to_none)
        g = self.grad_buffer()
        g.zero_()

    def _apply(self, fn, recurse=True):
        out = super()._apply(fn, recurse)
        self._flat = None
        self._flat_buf = None
        self._gflat = None
        self._flatten()
        return out

    def __deepcopy__(self, memo):
        import copy
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            if k in ("_flat", "_flat_buf", "_gflat"):
                new.__dict__[k] = None
            else:
                new.__dict__[k] = copy.deepcopy(v, memo)
        new._flatten()
        return new


def _param_list(module: nn.Module) -> List[nn.Parameter]:
    return list(module.parameters())


def _check_input(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise _lib.WggError(f"{name} must be a CUDA tensor: wordgesture-gan_b200 has no CPU execution path")
    if t.dtype != torch.float32:
        raise _lib.WggError(f"{name} must be float32 (the reference path is fp32, SURVEY.md section 5), got {t.dtype}")
    return t.contiguous()


def _linear_init_(weight: torch.Tensor, bias: torch.Tensor) -> None:
    """nn.Linear / nn.Conv1d default init (kaiming-uniform a=sqrt(5), then bias ~ U(+-1/sqrt(fan_in)))."""
    nn.init.kaiming_uniform_(weight, a=math.sqrt(5))
    fan_in = weight[0].numel()
    bound = 1 / math.sqrt(fan_in) if fan_in > 0 else 0
    nn.init.uniform_(bias, -bound, bound)


class _Holder(nn.Module):
    """Pure parameter container (never called)."""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter holder modules are not callable; call the parent module")


class _Dense(_Holder):
    def __init__(self, in_features: int, out_features: int):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(out_features, in_features))
        self.bias = nn.Parameter(torch.empty(out_features))
        _linear_init_(self.weight.data, self.bias.data)


class _SNLayer(_Holder):
    """Spectral-normalised Linear/Conv1d parameters with torch.nn.utils.spectral_norm's registration order:
    ``bias`` first, then ``weight_orig``; buffers ``weight_u`` (out) and ``weight_v`` (in*k)
    (torch/nn/utils/spectral_norm.py:141-176)."""

    def __init__(self, weight_shape: Tuple[int, ...]):
        super().__init__()
        weight = torch.empty(*weight_shape)
        bias = torch.empty(weight_shape[0])
        _linear_init_(weight, bias)
        h = weight_shape[0]
        w = weight.numel() // h
        u = nn.functional.normalize(weight.new_empty(h).normal_(0, 1), dim=0, eps=1e-12)
        v = nn.functional.normalize(weight.new_empty(w).normal_(0, 1), dim=0, eps=1e-12)
        self.bias = nn.Parameter(bias)
        self.weight_orig = nn.Parameter(weight)
        self.register_buffer("weight_u", u)
        self.register_buffer("weight_v", v)


class _Stack(_Holder):
    """Container whose children carry explicit names (mirrors nn.Sequential / nn.ModuleList indices)."""

    def __init__(self, named_children):
        super().__init__()
        for name, child in named_children:
            self.add_module(str(name), child)


class _LSTMParams(_Holder):
    """nn.LSTM(batch_first, bidirectional) parameter set with nn.LSTM's names, order and init."""

    def __init__(self, input_size: int, hidden_size: int, num_layers: int):
        super().__init__()
        self.input_size, self.hidden_size, self.num_layers = input_size, hidden_size, num_layers
        for layer in range(num_layers):
            in_l = input_size if layer == 0 else 2 * hidden_size
            for sfx in ("", "_reverse"):
                self.register_parameter(f"weight_ih_l{layer}{sfx}", nn.Parameter(torch.empty(4 * hidden_size, in_l)))
                self.register_parameter(f"weight_hh_l{layer}{sfx}", nn.Parameter(torch.empty(4 * hidden_size, hidden_size)))
                self.register_parameter(f"bias_ih_l{layer}{sfx}", nn.Parameter(torch.empty(4 * hidden_size)))
                self.register_parameter(f"bias_hh_l{layer}{sfx}", nn.Parameter(torch.empty(4 * hidden_size)))
        stdv = 1.0 / math.sqrt(hidden_size) if hidden_size > 0 else 0
        for p in self.parameters():
            nn.init.uniform_(p.data, -stdv, stdv)


# ------------------------------------------------------------------------------------------------
# Generator
# ------------------------------------------------------------------------------------------------
class _GeneratorFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, prototype, z, *params):
        lib = _lib.lib()
        dev = prototype.device
        c = _lib.ctx(dev)
        cfg = _lib.c_cfg(module.config)
        flat = module.flat_params()
        B = prototype.shape[0]
        need_grad = any(ctx.needs_input_grad[2:])
        out = torch.empty(B, module.config.seq_length, module.config.input_dim, dtype=torch.float32, device=dev)
        stash = None
        if need_grad:
            stash = torch.empty(lib.wgg_generator_stash_floats(cfg, B), dtype=torch.float32, device=dev)
        nws = lib.wgg_generator_workspace_floats(cfg, B, 0)
        ws = _lib.workspace(dev, nws)
        _lib.check(lib.wgg_generator_forward(c, cfg, _lib.ptr(flat), _lib.ptr(prototype), _lib.ptr(z), B,
                                             _lib.ptr(out), _lib.ptr(stash), _lib.ptr(ws),
                                             ws.numel() if ws is not None else 0, _lib.stream(dev)), c)
        ctx.module = module
        ctx.B = B
        ctx.stash = stash
        ctx.math_mode = _lib.get_math_mode()
        ctx.save_for_backward(out)
        return out

    @staticmethod
    def backward(ctx, dout):
        module = ctx.module
        lib = _lib.lib()
        (out,) = ctx.saved_tensors
        dev = out.device
        c = _lib.ctx(dev)
        cfg = _lib.c_cfg(module.config)
        flat = module.flat_params()
        if ctx.stash is None:
            raise _lib.WggError("generator backward called twice (the activation stash is consumed by backward)")
        if ctx.math_mode != _lib.get_math_mode():
            raise _lib.WggError("math mode changed between a generator forward and its backward")
        B = ctx.B
        dflat = module.grad_buffer()   # accumulated in place; the parameter inputs get None (see FlatModule.grad_buffer)
        dz = torch.empty(B, module.config.latent_dim, dtype=torch.float32, device=dev) if ctx.needs_input_grad[2] else None
        nws = lib.wgg_generator_workspace_floats(cfg, B, 1)
        ws = _lib.workspace(dev, nws)
        _lib.check(lib.wgg_generator_backward(c, cfg, _lib.ptr(flat), B, _lib.ptr(ctx.stash), _lib.ptr(out),
                                              _lib.ptr(dout.contiguous()), _lib.ptr(dflat), _lib.ptr(dz),
                                              _lib.ptr(ws), ws.numel(), _lib.stream(dev)), c)
        ctx.stash = None
        return (None, None, dz) + (None,) * (len(ctx.needs_input_grad) - 3)


class Generator(FlatModule):
    """Word-prototype-conditioned BiLSTM generator (reference: src/gan/models.py:89-165)."""

    def __init__(self, config: ModelConfig = DEFAULT_MODEL_CONFIG):
        super().__init__()
        self.config = config
        proto_dim = config.input_dim if config.prototype_has_time else 2
        self.lstm = _LSTMParams(proto_dim + config.latent_dim, config.gen_hidden_dim, config.gen_num_layers)
        self.output_layer = _Dense(config.gen_hidden_dim * 2, config.input_dim)
        self._flatten()

    def forward(self, prototype: torch.Tensor, z: torch.Tensor) -> torch.Tensor:
        prototype = _check_input(prototype, "prototype")
        z = _check_input(z, "z")
        if prototype.dim() != 3 or prototype.shape[1] != self.config.seq_length or prototype.shape[2] != self.config.input_dim:
            raise ValueError(f"prototype must be (B, {self.config.seq_length}, {self.config.input_dim}), got {tuple(prototype.shape)}")
        if z.shape != (prototype.shape[0], self.config.latent_dim):
            raise ValueError(f"z must be (B, {self.config.latent_dim}), got {tuple(z.shape)}")
        params = _param_list(self)
        if not torch.is_grad_enabled():
            return _GeneratorFn.apply(self, prototype, z.detach(), *[p.detach() for p in params])
        return _GeneratorFn.apply(self, prototype, z, *params)


# ------------------------------------------------------------------------------------------------
# Variational encoder
# ------------------------------------------------------------------------------------------------
class _EncoderFn(torch.autograd.Function):
    """(z, mu, log_var, kld): the fourth output is the batch-mean KL term of (mu, log_var), produced by the same fused
    kernel as the latent heads and the reparameterisation (csrc/encoder.cu: enc_head_fused_kernel)."""

    @staticmethod
    def forward(ctx, module, x, eps, *params):
        lib = _lib.lib()
        dev = x.device
        c = _lib.ctx(dev)
        cfg = _lib.c_cfg(module.config)
        flat = module.flat_params()
        B = x.shape[0]
        Z = module.config.latent_dim
        z = torch.empty(B, Z, dtype=torch.float32, device=dev)
        mu = torch.empty_like(z)
        log_var = torch.empty_like(z)
        want_kl = bool(getattr(module, "_want_kl", False)) and B > 0
        kld = torch.empty((), dtype=torch.float32, device=dev) if want_kl else torch.zeros((), dtype=torch.float32, device=dev)
        stash = torch.empty(lib.wgg_encoder_stash_floats(cfg, B), dtype=torch.float32, device=dev)
        _lib.check(lib.wgg_encoder_forward_kl(c, cfg, _lib.ptr(flat), _lib.ptr(x), _lib.ptr(eps), B, _lib.ptr(z),
                                              _lib.ptr(mu), _lib.ptr(log_var), _lib.ptr(stash),
                                              _lib.ptr(kld) if want_kl else None, _lib.stream(dev)), c)
        ctx.module = module
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(x, eps, mu, log_var, stash)
        return z, mu, log_var, kld

    @staticmethod
    def backward(ctx, dz, dmu, dlv, dkl):
        module = ctx.module
        lib = _lib.lib()
        x, eps, mu, log_var, stash = ctx.saved_tensors
        dev = x.device
        c = _lib.ctx(dev)
        cfg = _lib.c_cfg(module.config)
        flat = module.flat_params()
        B = x.shape[0]
        dflat = module.grad_buffer()
        dx = torch.empty_like(x) if ctx.needs_input_grad[1] else None
        nws = lib.wgg_encoder_workspace_floats(cfg, B)
        ws = _lib.workspace(dev, nws)
        cont = lambda t: None if t is None else t.contiguous()
        _lib.check(lib.wgg_encoder_backward_kl(c, cfg, _lib.ptr(flat), _lib.ptr(x), _lib.ptr(eps), _lib.ptr(mu),
                                               _lib.ptr(log_var), B, _lib.ptr(stash), _lib.ptr(cont(dz)),
                                               _lib.ptr(cont(dmu)), _lib.ptr(cont(dlv)), _lib.ptr(cont(dkl)),
                                               _lib.ptr(dflat), _lib.ptr(dx), _lib.ptr(ws), ws.numel(),
                                               _lib.stream(dev)), c)
        return (None, dx, None) + (None,) * (len(ctx.needs_input_grad) - 3)


class VariationalEncoder(FlatModule):
    """MLP encoder to a Gaussian latent (reference: src/gan/models.py:18-86)."""

    def __init__(self, config: ModelConfig = DEFAULT_MODEL_CONFIG):
        super().__init__()
        self.config = config
        dims = [config.seq_length * config.input_dim] + list(config.enc_hidden_dims)
        # nn.Sequential(Linear, LeakyReLU, ...) numbering: Linear layers sit at even indices
        self.encoder = _Stack((2 * i, _Dense(dims[i], dims[i + 1])) for i in range(len(dims) - 1))
        self.fc_mu = _Dense(dims[-1], config.latent_dim)
        self.fc_log_var = _Dense(dims[-1], config.latent_dim)
        self._flatten()

    def _run(self, x: torch.Tensor, eps: torch.Tensor, want_kl: bool = False):
        params = _param_list(self)
        self._want_kl = want_kl  # read by _EncoderFn.forward: the KL finalisation launch is skipped when unused
        if not torch.is_grad_enabled():
            return _EncoderFn.apply(self, x.detach(), eps, *[p.detach() for p in params])
        return _EncoderFn.apply(self, x, eps, *params)

    def forward(self, x: torch.Tensor, eps: torch.Tensor = None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """Returns (z, mu, log_var).  ``eps`` (optional) injects the reparameterisation noise; by default it is
        drawn with ``torch.randn`` exactly where the reference draws it (models.py:85), keeping the RNG stream."""
        x = _check_input(x, "x")
        B = x.shape[0]
        if x.numel() != B * self.config.seq_length * self.config.input_dim:
            raise ValueError(f"x must be (B, {self.config.seq_length}, {self.config.input_dim}), got {tuple(x.shape)}")
        if eps is None:
            eps = torch.randn(B, self.config.latent_dim, dtype=torch.float32, device=x.device)
        else:
            eps = _check_input(eps, "eps")
        return self._run(x, eps)[:3]

    def forward_with_kl(self, x: torch.Tensor, eps: torch.Tensor = None):
        """(z, mu, log_var, kld) with kld = KLDivergenceLoss()(mu, log_var) (src/gan/losses.py:174-175) computed by the
        fused latent-head kernel - what cycle 2 needs (trainer.py:161-171) in one pass."""
        x = _check_input(x, "x")
        B = x.shape[0]
        if x.numel() != B * self.config.seq_length * self.config.input_dim:
            raise ValueError(f"x must be (B, {self.config.seq_length}, {self.config.input_dim}), got {tuple(x.shape)}")
        if eps is None:
            eps = torch.randn(B, self.config.latent_dim, dtype=torch.float32, device=x.device)
        else:
            eps = _check_input(eps, "eps")
        return self._run(x, eps, want_kl=True)

    def reparameterize(self, mu: torch.Tensor, log_var: torch.Tensor) -> torch.Tensor:
        """z = mu + eps * exp(0.5 log_var), eps ~ N(0, I)   (models.py:78-86)."""
        std = torch.exp(0.5 * log_var)
        return mu + torch.randn_like(std) * std


# ------------------------------------------------------------------------------------------------
# Discriminators
# ------------------------------------------------------------------------------------------------
class _DiscFn(torch.autograd.Function):
    """One discriminator call: spectral-norm power iteration (in place on the u/v buffers when the module is in
    training mode) + forward.  Returns (score, stash) - score is None-like (empty) for feature-only calls.
    ``stash`` holds every post-activation feature block in kernel layout; its gradient is accepted in the same
    layout, so the feature-matching loss can be fused without re-layouts."""

    @staticmethod
    def forward(ctx, module, with_score, x, *params):
        lib = _lib.lib()
        dev = x.device
        c = _lib.ctx(dev)
        cfg = _lib.c_cfg(module.config)
        flat = module.flat_params()
        uv = module.flat_buffers()
        B = x.shape[0]
        st = _lib.stream(dev)
        sn = torch.zeros(lib.wgg_disc_sn_floats(cfg), dtype=torch.float32, device=dev)
        _lib.check(lib.wgg_disc_spectral(c, cfg, _lib.ptr(flat), _lib.ptr(uv), int(module.training), int(with_score),
                                         _lib.ptr(sn), st), c)
        stash = torch.empty(lib.wgg_disc_stash_floats(cfg, B), dtype=torch.float32, device=dev)
        score = torch.empty(B, 1, dtype=torch.float32, device=dev) if with_score else None
        _lib.check(lib.wgg_disc_forward(c, cfg, _lib.ptr(flat), _lib.ptr(sn), _lib.ptr(x), B, _lib.ptr(score),
                                        _lib.ptr(stash), st), c)
        ctx.module = module
        ctx.with_score = with_score
        ctx.math_mode = _lib.get_math_mode()
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(x, sn, stash)
        if with_score:
            return score, stash
        return stash.new_empty(0), stash

    @staticmethod
    def backward(ctx, dscore, dstash):
        module = ctx.module
        lib = _lib.lib()
        x, sn, stash = ctx.saved_tensors
        dev = x.device
        c = _lib.ctx(dev)
        cfg = _lib.c_cfg(module.config)
        flat = module.flat_params()
        B = x.shape[0]
        if not ctx.with_score:
            dscore = None
        if ctx.math_mode != _lib.get_math_mode():
            raise _lib.WggError("math mode changed between a discriminator forward and its backward")
        n_in = len(ctx.needs_input_grad)
        if dscore is None and dstash is None:
            return (None,) * n_in
        want_params = any(ctx.needs_input_grad[3:]) and not _lib.SKIP_DISC_WEIGHT_GRADS
        dflat = module.grad_buffer() if want_params else None
        dx = torch.empty_like(x) if ctx.needs_input_grad[2] else None
        nws = lib.wgg_disc_workspace_floats(cfg, B)
        ws = _lib.workspace(dev, nws)
        cont = lambda t: None if t is None else t.contiguous()
        _lib.check(lib.wgg_disc_backward(c, cfg, _lib.ptr(flat), _lib.ptr(sn), _lib.ptr(x), B, _lib.ptr(stash),
                                         _lib.ptr(cont(dscore)), _lib.ptr(cont(dstash)), _lib.ptr(dflat), _lib.ptr(dx),
                                         _lib.ptr(ws), ws.numel(), _lib.stream(dev)), c)
        return (None, None, dx) + (None,) * (n_in - 3)


class _FeatureViewFn(torch.autograd.Function):
    """stash -> feature k in the layout get_all_features documents: (B, n_k), conv features flattened
    channel-major as ``h.view(B, -1)`` of a (B, C, T) tensor (models.py:339)."""

    @staticmethod
    def forward(ctx, stash, config, off, B, T, C, conv):
        ctx.meta = (config, off, B, T, C, conv, stash.numel(), _lib.get_math_mode())
        block = stash[off:off + B * T * C]
        if not conv:
            return block.view(B, T * C).clone()
        out = torch.empty(B, C * T, dtype=torch.float32, device=stash.device)
        c = _lib.ctx(stash.device)
        _lib.check(_lib.lib().wgg_disc_feature_convert(c, _lib.c_cfg(config), _lib.ptr(block), _lib.ptr(out), B, T, C, 0,
                                                       _lib.stream(stash.device)), c)
        return out

    @staticmethod
    def backward(ctx, g):
        config, off, B, T, C, conv, n, mode = ctx.meta
        if mode != _lib.get_math_mode():
            raise _lib.WggError("math mode changed between a discriminator forward and its backward")
        dst = torch.zeros(n, dtype=torch.float32, device=g.device)
        g = g.contiguous()
        if not conv:
            dst[off:off + B * T * C].copy_(g.reshape(-1))
        else:
            c = _lib.ctx(g.device)
            _lib.check(_lib.lib().wgg_disc_feature_convert(c, _lib.c_cfg(config), _lib.ptr(g),
                                                           _lib.ptr(dst[off:off + B * T * C]), B, T, C, 1,
                                                           _lib.stream(g.device)), c)
        return dst, None, None, None, None, None, None


class _DiscBase(FlatModule):
    config: ModelConfig

    def _call(self, x: torch.Tensor, with_score: bool):
        x = _check_input(x, "x")
        B = x.shape[0]
        if x.numel() != B * self.config.seq_length * self.config.input_dim:
            raise ValueError(f"x must be (B, {self.config.seq_length}, {self.config.input_dim}), got {tuple(x.shape)}")
        x = x.view(B, self.config.seq_length, self.config.input_dim)
        params = _param_list(self)
        if not torch.is_grad_enabled():
            return _DiscFn.apply(self, with_score, x.detach(), *[p.detach() for p in params])
        return _DiscFn.apply(self, with_score, x, *params)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """(B, T, 3) -> critic score (B, 1)."""
        return self._call(x, True)[0]

    def features_stash(self, x: torch.Tensor) -> torch.Tensor:
        """Feature-only call (advances every power iteration except output_layer's, exactly like
        get_all_features) returning the raw stash; pair with losses.feature_matching_from_stash."""
        return self._call(x, False)[1]

    def _conv_feature_channels(self, k: int) -> int:
        return 0

    def get_all_features(self, x: torch.Tensor) -> list:
        stash = self.features_stash(x)
        lib = _lib.lib()
        cfg = _lib.c_cfg(self.config)
        B = x.shape[0]
        feats = []
        for k in range(lib.wgg_disc_num_features(cfg)):
            off = lib.wgg_disc_feature_offset(cfg, B, k)
            width = lib.wgg_disc_feature_width(cfg, k)
            C = self._conv_feature_channels(k)
            if C:
                feats.append(_FeatureViewFn.apply(stash, self.config, off, B, width // C, C, True))
            else:
                feats.append(_FeatureViewFn.apply(stash, self.config, off, B, 1, width, False))
        return feats


class Discriminator(_DiscBase):
    """Spectral-normalised MLP critic (reference: src/gan/models.py:168-243)."""

    def __init__(self, config: ModelConfig = DEFAULT_MODEL_CONFIG):
        super().__init__()
        self.config = config
        dims = [config.seq_length * config.input_dim] + list(config.disc_hidden_dims)
        self.layers = _Stack((i, _SNLayer((dims[i + 1], dims[i]))) for i in range(len(dims) - 1))
        self.output_layer = _SNLayer((1, dims[-1]))
        self._flatten()


class TemporalDiscriminator(_DiscBase):
    """Spectral-normalised Conv1D critic, the reference's default (src/gan/models.py:246-353)."""

    def __init__(self, config: ModelConfig = DEFAULT_MODEL_CONFIG):
        super().__init__()
        self.config = config
        self.temporal_conv = _Stack((
            (0, _SNLayer((64, config.input_dim, 5))),
            (2, _SNLayer((64, 64, 5))),
            (4, _SNLayer((32, 64, 3))),
        ))
        self.mlp = _Stack(((0, _SNLayer((128, 32 * 8))), (2, _SNLayer((64, 128)))))
        self.output_layer = _SNLayer((1, 64))
        self._flatten()

    def _conv_feature_channels(self, k: int) -> int:
        return (64, 64, 32, 0, 0)[k]
