"""ctypes binding of libwgg_sm100.so (the C ABI declared in include/wgg.h).

There is NO fallback: if the shared library is missing or the device is not a CUDA sm_100 GPU every
compute entry point raises.  PyTorch is used only for device memory, streams and autograd plumbing.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int32, c_int64, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libwgg_sm100.so")
MAX_HIDDEN = 8


class WggError(RuntimeError):
    pass


class CModelCfg(ctypes.Structure):
    _fields_ = [
        ("seq_length", c_int32), ("input_dim", c_int32), ("latent_dim", c_int32),
        ("gen_hidden_dim", c_int32), ("gen_num_layers", c_int32),
        ("prototype_has_time", c_int32), ("use_temporal_disc", c_int32),
        ("n_enc_hidden", c_int32), ("enc_hidden_dims", c_int32 * MAX_HIDDEN),
        ("n_disc_hidden", c_int32), ("disc_hidden_dims", c_int32 * MAX_HIDDEN),
    ]


_P = c_void_p
_CFG = POINTER(CModelCfg)
_SIG = {
    "wgg_abi_version": (c_int, []),
    "wgg_create": (c_int, [POINTER(c_void_p), c_int]),
    "wgg_destroy": (None, [_P]),
    "wgg_last_error": (c_char_p, [_P]),
    "wgg_launch_count": (c_int64, [_P]),
    "wgg_set_math_mode": (c_int, [_P, c_int]),
    "wgg_set_lane": (c_int, [_P, c_int]),
    "wgg_async_error": (c_int, [_P, POINTER(c_int)]),
    "wgg_profile_report": (c_int, [_P, c_char_p, c_int64]),
    "wgg_profile_enable": (c_int, [_P, c_char_p]),
    "wgg_profile_read": (c_int, [_P, POINTER(ctypes.c_double), POINTER(c_int64), POINTER(ctypes.c_double),
                                 POINTER(ctypes.c_double)]),
    "wgg_generator_param_floats": (c_int64, [_CFG]),
    "wgg_generator_stash_floats": (c_int64, [_CFG, c_int64]),
    "wgg_generator_workspace_floats": (c_int64, [_CFG, c_int64, c_int]),
    "wgg_generator_forward": (c_int, [_P, _CFG, _P, _P, _P, c_int64, _P, _P, _P, c_int64, _P]),
    "wgg_generator_backward": (c_int, [_P, _CFG, _P, c_int64, _P, _P, _P, _P, _P, _P, c_int64, _P]),
    "wgg_encoder_param_floats": (c_int64, [_CFG]),
    "wgg_encoder_stash_floats": (c_int64, [_CFG, c_int64]),
    "wgg_encoder_workspace_floats": (c_int64, [_CFG, c_int64]),
    "wgg_encoder_forward": (c_int, [_P, _CFG, _P, _P, _P, c_int64, _P, _P, _P, _P, _P]),
    "wgg_encoder_backward": (c_int, [_P, _CFG, _P, _P, _P, _P, c_int64, _P, _P, _P, _P, _P, _P, _P, c_int64, _P]),
    "wgg_encoder_forward_kl": (c_int, [_P, _CFG, _P, _P, _P, c_int64, _P, _P, _P, _P, _P, _P]),
    "wgg_encoder_backward_kl": (c_int, [_P, _CFG, _P, _P, _P, _P, _P, c_int64, _P, _P, _P, _P, _P, _P, _P, _P, c_int64, _P]),
    "wgg_disc_param_floats": (c_int64, [_CFG]),
    "wgg_disc_uv_floats": (c_int64, [_CFG]),
    "wgg_disc_sn_floats": (c_int64, [_CFG]),
    "wgg_disc_stash_floats": (c_int64, [_CFG, c_int64]),
    "wgg_disc_workspace_floats": (c_int64, [_CFG, c_int64]),
    "wgg_disc_num_features": (c_int32, [_CFG]),
    "wgg_disc_feature_offset": (c_int64, [_CFG, c_int64, c_int32]),
    "wgg_disc_feature_width": (c_int32, [_CFG, c_int32]),
    "wgg_disc_spectral": (c_int, [_P, _CFG, _P, _P, c_int, c_int, _P, _P]),
    "wgg_disc_forward": (c_int, [_P, _CFG, _P, _P, _P, c_int64, _P, _P, _P]),
    "wgg_disc_backward": (c_int, [_P, _CFG, _P, _P, _P, c_int64, _P, _P, _P, _P, _P, _P, c_int64, _P]),
    "wgg_transpose_tc": (c_int, [_P, _P, _P, c_int64, c_int32, c_int32, _P]),
    "wgg_disc_feature_convert": (c_int, [_P, _CFG, _P, _P, c_int64, c_int32, c_int32, c_int, _P]),
    "wgg_mean": (c_int, [_P, _P, c_int64, c_float, c_int, _P, _P]),
    "wgg_mean_backward": (c_int, [_P, _P, c_float, c_int64, _P, _P]),
    "wgg_l1_mean": (c_int, [_P, _P, _P, c_int64, c_float, c_int, _P, _P]),
    "wgg_l1_mean_backward": (c_int, [_P, _P, _P, _P, c_float, c_int64, c_int, _P, _P]),
    "wgg_feature_matching": (c_int, [_P, _CFG, _P, _P, c_int64, c_float, c_int, _P, _P]),
    "wgg_feature_matching_backward": (c_int, [_P, _CFG, _P, _P, _P, c_float, c_int64, _P, _P]),
    "wgg_kl": (c_int, [_P, _P, _P, c_int64, c_int32, c_float, c_int, _P, _P]),
    "wgg_kl_backward": (c_int, [_P, _P, _P, _P, c_float, c_int64, c_int32, _P, _P, _P]),
    "wgg_clip_adam_workspace_floats": (c_int64, []),
    "wgg_clip_adam": (c_int, [_P, _P, _P, _P, _P, c_int64, c_float, c_float, c_float, c_float, c_int64, c_float,
                              _P, _P, _P]),
    "wgg_clip_adam_dev": (c_int, [_P, _P, _P, _P, _P, c_int64, _P, c_float, c_float, c_float, _P, c_float, _P, _P, _P]),
    "wgg_linear": (c_int, [_P, _P, _P, _P, _P, c_int64, c_int32, c_int32, c_int, _P]),
    "wgg_eval_cdist": (c_int, [_P, _P, c_int64, _P, c_int64, c_int32, _P, _P]),
    "wgg_eval_row_kth": (c_int, [_P, _P, c_int64, c_int64, c_int32, _P, _P]),
    "wgg_eval_precision_recall": (c_int, [_P, _P, c_int64, c_int64, _P, _P, _P, _P, _P]),
    "wgg_eval_jerk": (c_int, [_P, _P, c_int64, c_int32, c_int32, _P, _P, _P, _P]),
    "wgg_eval_dynamics": (c_int, [_P, _P, _P, c_int64, c_int32, c_int32, _P, _P, _P]),
    "wgg_p2p_flag_words": (c_int64, []),
    "wgg_p2p_alloc": (c_int, [_P, c_int64, POINTER(c_void_p), ctypes.c_char_p]),
    "wgg_p2p_open": (c_int, [_P, ctypes.c_char_p, POINTER(c_void_p)]),
    "wgg_p2p_close": (c_int, [_P, _P, c_int]),
    "wgg_p2p_allreduce_avg": (c_int, [_P, _P, _P, c_int, c_int, c_int64, _P, _P, _P, _P]),
    "wgg_word_prototypes": (c_int, [_P, _P, _P, c_int64, c_int32, c_int32, _P, _P]),
    "wgg_minimum_jerk": (c_int, [_P, _P, _P, c_int64, c_int32, _P, _P, c_int, c_int32, _P, _P]),
}
EXPORTED_SYMBOLS = tuple(_SIG)

_lib = None
_ctx = {}
_ws = {}
_cfg_cache = {}


def lib():
    """Loads the shared library (no CUDA call is made here)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise WggError(
                f"{LIB_PATH} is missing - build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU or PyTorch fallback)")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIG.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        if handle.wgg_abi_version() != 1:
            raise WggError("libwgg_sm100.so ABI version mismatch")
        _lib = handle
    return _lib


def ctx(device: torch.device):
    """One wgg_ctx per CUDA device of this process."""
    device = torch.device(device)
    if device.type != "cuda":
        raise WggError("wordgesture-gan_b200 runs on CUDA sm_100a devices only (no CPU path); got " + str(device))
    idx = device.index if device.index is not None else torch.cuda.current_device()
    c = _ctx.get(idx)
    if c is None:
        out = c_void_p()
        rc = lib().wgg_create(ctypes.byref(out), idx)
        if rc != 0:
            raise WggError(f"wgg_create(device={idx}) failed with code {rc} (needs an sm_100 GPU)")
        c = _ctx[idx] = out
        if _math_mode:
            check(lib().wgg_set_math_mode(c, _math_mode), c)
    return c


def check(rc: int, c) -> None:
    if rc != 0:
        msg = lib().wgg_last_error(c)
        raise WggError(f"libwgg_sm100 error {rc}: {msg.decode() if msg else ''}")


def launch_count(device) -> int:
    return int(lib().wgg_launch_count(ctx(device)))


def c_cfg(mc) -> CModelCfg:
    key = (mc.seq_length, mc.input_dim, mc.latent_dim, mc.gen_hidden_dim, mc.gen_num_layers,
           bool(mc.prototype_has_time), bool(mc.use_temporal_disc), tuple(mc.enc_hidden_dims),
           tuple(mc.disc_hidden_dims))
    c = _cfg_cache.get(key)
    if c is None:
        if len(key[7]) > MAX_HIDDEN or len(key[8]) > MAX_HIDDEN:
            raise WggError("too many hidden layers")
        c = CModelCfg()
        c.seq_length, c.input_dim, c.latent_dim, c.gen_hidden_dim, c.gen_num_layers = key[:5]
        c.prototype_has_time, c.use_temporal_disc = int(key[5]), int(key[6])
        c.n_enc_hidden = len(key[7])
        for i, v in enumerate(key[7]):
            c.enc_hidden_dims[i] = v
        c.n_disc_hidden = len(key[8])
        for i, v in enumerate(key[8]):
            c.disc_hidden_dims[i] = v
        _cfg_cache[key] = c
    return c


def ptr(t):
    """Raw device pointer of an fp32 contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if t.dtype != torch.float32 or not t.is_cuda or not t.is_contiguous():
        raise WggError(f"expected a contiguous float32 CUDA tensor, got {t.dtype} {t.device} "
                       f"contiguous={t.is_contiguous()}")
    return t.data_ptr()


def stream(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


_lane = 0


class lane:
    """Context manager for calls issued on the SECOND of two concurrently driven streams (train_step runs the two
    critic chains side by side): they get their own scratch workspace and their own half of the library's
    reduction-scratch ring (wgg_set_lane)."""

    def __init__(self, device, k: int = 1):
        self.device, self.k = device, k

    def __enter__(self):
        global _lane
        self.prev = _lane
        _lane = self.k
        c = ctx(self.device)
        check(lib().wgg_set_lane(c, self.k), c)
        return self

    def __exit__(self, *exc):
        global _lane
        _lane = self.prev
        c = ctx(self.device)
        check(lib().wgg_set_lane(c, self.prev), c)
        return False


_scope = None  # the active ScratchScope (None: the process-wide grow-only buffers in _ws)


class ScratchScope:
    """A private set of per-(device, lane) scratch buffers.  A captured CUDA graph bakes the addresses of the scratch
    its kernels use into the graph, so the graph must OWN that scratch: GraphedTrainStep sizes a scope during its
    warm-up, captures under the same scope (frozen: growing during capture would leave earlier captured launches
    pointing at a freed buffer) and keeps the scope alive for as long as the graph lives.  Eager calls made outside
    the scope use (and may regrow) the process-wide buffers without touching the graph's."""

    def __init__(self):
        self.buffers = {}
        self.frozen = False
        self._prev = None

    def __enter__(self):
        global _scope
        self._prev = _scope
        _scope = self
        return self

    def __exit__(self, *exc):
        global _scope
        _scope = self._prev
        return False


def workspace(device, nfloats: int) -> torch.Tensor:
    """Grow-only scratch buffer per (device, lane).  All calls of one lane are issued on one stream at a time, so a
    single buffer is shared by every entry point of that lane.  Inside a ScratchScope the scope's own buffers are
    used."""
    device = torch.device(device)
    idx = device.index if device.index is not None else torch.cuda.current_device()
    key = (idx, _lane)
    store = _ws if _scope is None else _scope.buffers
    w = store.get(key)
    if w is None or w.numel() < nfloats:
        if _scope is not None and _scope.frozen:
            raise WggError(f"scratch request of {nfloats} floats exceeds the {0 if w is None else w.numel()} floats sized "
                           "by the warm-up of a captured graph (the warm-up must run the same step as the capture)")
        if torch.cuda.is_current_stream_capturing():
            raise WggError("scratch buffers cannot grow during CUDA graph capture; warm the step up first")
        store[key] = None
        w = store[key] = torch.empty(max(int(nfloats), 1 << 20), dtype=torch.float32, device=device)
    return w


def profile_enable(device, kernel_substr):
    c = ctx(device)
    check(lib().wgg_profile_enable(c, kernel_substr.encode() if kernel_substr else None), c)


def profile_read(device):
    c = ctx(device)
    ms, n, fl, by = ctypes.c_double(), c_int64(), ctypes.c_double(), ctypes.c_double()
    check(lib().wgg_profile_read(c, ctypes.byref(ms), ctypes.byref(n), ctypes.byref(fl), ctypes.byref(by)), c)
    return dict(ms=ms.value, launches=n.value, flops=fl.value, bytes=by.value)


_math_mode = 0


def set_math_mode(mode, device=None) -> None:
    """"fp32" / 0: every contraction in fp32 FMA (bit-for-bit the most faithful, used by the tight parity tests).
    "tf32" / 1: LSTM and conv contractions on TF32 tensor cores (tcgen05 kernels) with fp32 accumulation - the
    numerics of the reference's own CUDA path (cuDNN allows TF32 by default); nn.Linear layers
    stay fp32.  "tf32x3" / 2: as "tf32" but conv contractions in error-compensated 3xTF32 (fp32-grade)."""
    global _math_mode
    m = {"fp32": 0, "tf32": 1, "tf32x3": 2}.get(mode, mode)
    if m not in (0, 1, 2):
        raise ValueError(f"unknown math mode {mode!r}")
    _math_mode = m
    for idx, c in _ctx.items():
        check(lib().wgg_set_math_mode(c, m), c)


def get_math_mode() -> str:
    return ("fp32", "tf32", "tf32x3")[_math_mode]


def async_error(device) -> int:
    """Synchronising debug query of the persistent kernels' pipeline-timeout word (0 = healthy)."""
    c = ctx(device)
    code = c_int(0)
    check(lib().wgg_async_error(c, ctypes.byref(code)), c)
    return code.value


def profile_report(device):
    c = ctx(device)
    buf = ctypes.create_string_buffer(16384)
    check(lib().wgg_profile_report(c, buf, 16384), c)
    rows = []
    for line in buf.value.decode().splitlines():
        tag, n, ms, gf = line.rsplit(" ", 3)
        rows.append(dict(tag=tag, launches=int(n), ms=float(ms), gflop=float(gf)))
    return rows


# Set by train_step while the generator/encoder step back-propagates through the discriminators: their weight
# gradients are discarded by the reference (zero_grad before the next critic backward, utils.py:75,96), so the
# discriminator backward may skip computing them.
SKIP_DISC_WEIGHT_GRADS = False
