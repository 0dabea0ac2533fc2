"""numpy restatement of the WordGesture-GAN training-step hot path (CPU oracle).

TEST INFRASTRUCTURE - see ``oracle/__init__.py``.  Every function cites the
reference file:line (relative to /root/reference) whose arithmetic it restates.
All maths is explicit (forward AND hand-derived backward), in ``dtype``
(float64 by default, float32 for the CPU-baseline timing leg).

Parity status: PINNED against the reference itself - ``oracle/make_golden.py``
runs the unmodified reference and this file on the same inputs, asserts
agreement, and writes ``tests/golden/*.npz``; ``tests/test_oracle_golden.py``
re-checks this file against those fixtures wherever the tests run.

State layout: plain dicts ``name -> ndarray`` using the reference's
``state_dict`` keys, one dict per module.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

Array = np.ndarray


# ----------------------------------------------------------------------------------------------
# configuration mirrors (src/shared/config.py:11-58) - same field names and defaults
# ----------------------------------------------------------------------------------------------
@dataclass
class ModelCfg:
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
class TrainCfg:
    learning_rate: float = 0.0002
    n_critic: int = 5
    lambda_feat: float = 1.0
    lambda_rec: float = 4.0
    lambda_lat: float = 0.5
    lambda_kld: float = 0.02
    betas: Tuple[float, float] = (0.5, 0.999)  # src/gan/trainer.py:60-79
    adam_eps: float = 1e-8


LEAK = 0.2  # nn.LeakyReLU(0.2) - src/gan/models.py:43,200,272


def _sigmoid(x: Array) -> Array:
    return 1.0 / (1.0 + np.exp(-x))


def leaky(x: Array) -> Array:
    return np.where(x > 0, x, LEAK * x)


def leaky_bwd(y: Array, dy: Array) -> Array:
    # torch leaky_relu_backward: grad where x > 0 else grad*slope (y has the sign of x)
    return np.where(y > 0, dy, LEAK * dy)


# ----------------------------------------------------------------------------------------------
# LSTM (nn.LSTM semantics; gate order i,f,g,o; src/gan/models.py:114-120,160)
# ----------------------------------------------------------------------------------------------
def lstm_dir_fwd(x: Array, w_ih: Array, w_hh: Array, b_ih: Array, b_hh: Array, reverse: bool):
    """One direction of one layer.  x (B,T,I) -> h (B,T,H) plus the stash for backward."""
    B, T, _ = x.shape
    H = w_hh.shape[1]
    dt = x.dtype
    hs = np.zeros((B, T, H), dt)
    gates = np.zeros((B, T, 4 * H), dt)  # post-activation i,f,g,o
    cs = np.zeros((B, T, H), dt)
    h = np.zeros((B, H), dt)
    c = np.zeros((B, H), dt)
    bias = b_ih + b_hh
    order = range(T - 1, -1, -1) if reverse else range(T)
    for t in order:
        a = x[:, t, :] @ w_ih.T + h @ w_hh.T + bias
        i = _sigmoid(a[:, 0:H])
        f = _sigmoid(a[:, H:2 * H])
        g = np.tanh(a[:, 2 * H:3 * H])
        o = _sigmoid(a[:, 3 * H:4 * H])
        c = f * c + i * g
        h = o * np.tanh(c)
        gates[:, t, 0:H] = i
        gates[:, t, H:2 * H] = f
        gates[:, t, 2 * H:3 * H] = g
        gates[:, t, 3 * H:4 * H] = o
        cs[:, t] = c
        hs[:, t] = h
    return hs, (gates, cs)


def lstm_dir_bwd(x: Array, hs: Array, stash, w_ih: Array, w_hh: Array, dh_out: Array, reverse: bool):
    """BPTT for one direction.  Returns dx, dW_ih, dW_hh, db (db_ih == db_hh == db)."""
    gates, cs = stash
    B, T, _ = x.shape
    H = w_hh.shape[1]
    dt = x.dtype
    dx = np.zeros_like(x)
    dW_ih = np.zeros_like(w_ih)
    dW_hh = np.zeros_like(w_hh)
    db = np.zeros(4 * H, dt)
    dh_rec = np.zeros((B, H), dt)
    dc = np.zeros((B, H), dt)
    # walk time in the opposite order of the forward scan
    order = list(range(T - 1, -1, -1) if reverse else range(T))
    for idx in range(T - 1, -1, -1):
        t = order[idx]
        t_prev = order[idx - 1] if idx > 0 else None
        i = gates[:, t, 0:H]
        f = gates[:, t, H:2 * H]
        g = gates[:, t, 2 * H:3 * H]
        o = gates[:, t, 3 * H:4 * H]
        c = cs[:, t]
        c_prev = cs[:, t_prev] if t_prev is not None else np.zeros((B, H), dt)
        h_prev = hs[:, t_prev] if t_prev is not None else np.zeros((B, H), dt)
        tc = np.tanh(c)
        dh = dh_out[:, t] + dh_rec
        do = dh * tc
        dc = dc + dh * o * (1.0 - tc * tc)
        di = dc * g
        df = dc * c_prev
        dg = dc * i
        da = np.concatenate(
            [di * i * (1.0 - i), df * f * (1.0 - f), dg * (1.0 - g * g), do * o * (1.0 - o)], axis=1
        )
        dc = dc * f
        dh_rec = da @ w_hh
        dx[:, t] = da @ w_ih
        dW_ih += da.T @ x[:, t]
        dW_hh += da.T @ h_prev
        db += da.sum(axis=0)
    return dx, dW_ih, dW_hh, db


def _lstm_keys(layer: int, reverse: bool):
    sfx = f"_l{layer}" + ("_reverse" if reverse else "")
    return (f"lstm.weight_ih{sfx}", f"lstm.weight_hh{sfx}", f"lstm.bias_ih{sfx}", f"lstm.bias_hh{sfx}")


# ----------------------------------------------------------------------------------------------
# Generator (src/gan/models.py:125-165)
# ----------------------------------------------------------------------------------------------
def generator_fwd(p: Dict[str, Array], cfg: ModelCfg, prototype: Array, z: Array, want_stash: bool = True):
    B, T = prototype.shape[:2]
    pd = cfg.input_dim if cfg.prototype_has_time else 2  # models.py:147-151
    x = np.concatenate([prototype[:, :, :pd], np.repeat(z[:, None, :], T, axis=1)], axis=-1)  # :154-157
    stash = {"inputs": [], "hs": [], "st": []}
    inp = x
    for layer in range(cfg.gen_num_layers):
        outs = []
        sts = []
        for rev in (False, True):
            k = _lstm_keys(layer, rev)
            hs, st = lstm_dir_fwd(inp, p[k[0]], p[k[1]], p[k[2]], p[k[3]], rev)
            outs.append(hs)
            sts.append(st)
        if want_stash:
            stash["inputs"].append(inp)
            stash["hs"].append(outs)
            stash["st"].append(sts)
        inp = np.concatenate(outs, axis=-1)
    y = np.tanh(inp @ p["output_layer.weight"].T + p["output_layer.bias"])  # models.py:163
    stash["last"] = inp
    stash["y"] = y
    stash["pd"] = pd
    return y, stash


def generator_bwd(p: Dict[str, Array], cfg: ModelCfg, stash, dy: Array):
    """Returns (param grads dict, dz (B,Z))."""
    H = cfg.gen_hidden_dim
    grads: Dict[str, Array] = {}
    y = stash["y"]
    last = stash["last"]
    dpre = dy * (1.0 - y * y)
    grads["output_layer.weight"] = np.einsum("btc,bth->ch", dpre, last)
    grads["output_layer.bias"] = dpre.sum(axis=(0, 1))
    dinp = dpre @ p["output_layer.weight"]
    for layer in range(cfg.gen_num_layers - 1, -1, -1):
        x = stash["inputs"][layer]
        dx_total = np.zeros_like(x)
        for d, rev in enumerate((False, True)):
            k = _lstm_keys(layer, rev)
            dx, dWi, dWh, db = lstm_dir_bwd(
                x, stash["hs"][layer][d], stash["st"][layer][d], p[k[0]], p[k[1]], dinp[:, :, d * H:(d + 1) * H], rev
            )
            grads[k[0]] = dWi
            grads[k[1]] = dWh
            grads[k[2]] = db
            grads[k[3]] = db.copy()
            dx_total += dx
        dinp = dx_total
    dz = dinp[:, :, stash["pd"]:].sum(axis=1)
    return grads, dz


# ----------------------------------------------------------------------------------------------
# Variational encoder (src/gan/models.py:52-86)
# ----------------------------------------------------------------------------------------------
def _enc_layer_ids(cfg: ModelCfg):
    return [2 * i for i in range(len(cfg.enc_hidden_dims))]  # nn.Sequential indices 0,2,4,6


def encoder_fwd(p: Dict[str, Array], cfg: ModelCfg, x: Array, eps: Array):
    B = x.shape[0]
    h = x.reshape(B, -1)  # models.py:64
    acts = [h]
    for i in _enc_layer_ids(cfg):
        h = leaky(h @ p[f"encoder.{i}.weight"].T + p[f"encoder.{i}.bias"])
        acts.append(h)
    mu = h @ p["fc_mu.weight"].T + p["fc_mu.bias"]
    log_var = h @ p["fc_log_var.weight"].T + p["fc_log_var.bias"]
    std = np.exp(0.5 * log_var)  # models.py:84
    z = mu + eps * std  # models.py:86
    return z, mu, log_var, {"acts": acts, "std": std, "eps": eps}


def encoder_bwd(p: Dict[str, Array], cfg: ModelCfg, stash, dz: Array, dmu: Array, dlog_var: Array):
    acts = stash["acts"]
    grads: Dict[str, Array] = {}
    dmu_t = dz + dmu
    dlv_t = dz * stash["eps"] * 0.5 * stash["std"] + dlog_var
    h = acts[-1]
    grads["fc_mu.weight"] = dmu_t.T @ h
    grads["fc_mu.bias"] = dmu_t.sum(0)
    grads["fc_log_var.weight"] = dlv_t.T @ h
    grads["fc_log_var.bias"] = dlv_t.sum(0)
    dh = dmu_t @ p["fc_mu.weight"] + dlv_t @ p["fc_log_var.weight"]
    ids = _enc_layer_ids(cfg)
    for n in range(len(ids) - 1, -1, -1):
        i = ids[n]
        dpre = leaky_bwd(acts[n + 1], dh)
        grads[f"encoder.{i}.weight"] = dpre.T @ acts[n]
        grads[f"encoder.{i}.bias"] = dpre.sum(0)
        dh = dpre @ p[f"encoder.{i}.weight"]
    return grads


# ----------------------------------------------------------------------------------------------
# Spectral norm (torch/nn/utils/spectral_norm.py:62-114): one power iteration per train-mode call
# ----------------------------------------------------------------------------------------------
SN_EPS = 1e-12


def _normalize(v: Array) -> Array:
    return v / max(float(np.sqrt((v * v).sum())), SN_EPS)


def sn_effective_weight(p: Dict[str, Array], prefix: str, training: bool):
    """Updates p[prefix.weight_u/_v] IN PLACE when training; returns (W_eff, sigma, u, v)."""
    w = p[prefix + ".weight_orig"]
    wm = w.reshape(w.shape[0], -1)
    u = p[prefix + ".weight_u"]
    v = p[prefix + ".weight_v"]
    if training:
        v = _normalize(wm.T @ u)
        u = _normalize(wm @ v)
        p[prefix + ".weight_v"] = v
        p[prefix + ".weight_u"] = u
    sigma = float(u @ (wm @ v))
    return w / sigma, sigma, u.copy(), v.copy()


def sn_weight_grad(w_orig: Array, g_eff: Array, sigma: float, u: Array, v: Array) -> Array:
    """dL/dW_orig given G = dL/dW_eff with u, v treated as constants."""
    inner = float((g_eff * w_orig).sum())
    return g_eff / sigma - (inner / (sigma * sigma)) * np.outer(u, v).reshape(w_orig.shape)


# ----------------------------------------------------------------------------------------------
# conv1d helpers (channel-first maths as in nn.Conv1d; x (B,Cin,T), w (Cout,Cin,K), padding=(K-1)//2)
# ----------------------------------------------------------------------------------------------
def conv1d_fwd(x: Array, w: Array, b: Array) -> Array:
    B, Cin, T = x.shape
    Cout, _, K = w.shape
    pad = (K - 1) // 2
    xp = np.zeros((B, Cin, T + 2 * pad), x.dtype)
    xp[:, :, pad:pad + T] = x
    out = np.zeros((B, Cout, T), x.dtype)
    for k in range(K):
        out += np.einsum("oc,bct->bot", w[:, :, k], xp[:, :, k:k + T])
    return out + b[None, :, None]


def conv1d_bwd(x: Array, w: Array, dout: Array):
    B, Cin, T = x.shape
    Cout, _, K = w.shape
    pad = (K - 1) // 2
    xp = np.zeros((B, Cin, T + 2 * pad), x.dtype)
    xp[:, :, pad:pad + T] = x
    dxp = np.zeros_like(xp)
    dw = np.zeros_like(w)
    for k in range(K):
        dw[:, :, k] = np.einsum("bot,bct->oc", dout, xp[:, :, k:k + T])
        dxp[:, :, k:k + T] += np.einsum("oc,bot->bct", w[:, :, k], dout)
    return dxp[:, :, pad:pad + T], dw, dout.sum(axis=(0, 2))


def adaptive_avg_pool_bins(T: int, n: int):
    """AdaptiveAvgPool1d bins: start=floor(i*T/n), end=ceil((i+1)*T/n)."""
    return [(int(math.floor(i * T / n)), int(math.ceil((i + 1) * T / n))) for i in range(n)]


# ----------------------------------------------------------------------------------------------
# TemporalDiscriminator (src/gan/models.py:261-353)
# ----------------------------------------------------------------------------------------------
TD_CONV = ("temporal_conv.0", "temporal_conv.2", "temporal_conv.4")
TD_MLP = ("mlp.0", "mlp.2")
TD_OUT = "output_layer"
POOL_BINS = 8


def tdisc_fwd(p: Dict[str, Array], cfg: ModelCfg, x: Array, training: bool = True, features_only: bool = False):
    """forward() (models.py:293-317) or get_all_features() (:319-353, which skips output_layer and so
    does NOT advance output_layer's power iteration).  Mutates the u/v buffers in p when training."""
    B = x.shape[0]
    h = np.transpose(x, (0, 2, 1))  # (B,3,T)
    st = {"sn": {}, "conv_in": [], "conv_out": [], "mlp_in": [], "mlp_out": []}
    feats = []
    for name in TD_CONV:
        w_eff, sigma, u, v = sn_effective_weight(p, name, training)
        st["sn"][name] = (sigma, u, v)
        st["conv_in"].append(h)
        h = leaky(conv1d_fwd(h, w_eff, p[name + ".bias"]))
        st["conv_out"].append(h)
        feats.append(h.reshape(B, -1))
    T = h.shape[2]
    bins = adaptive_avg_pool_bins(T, POOL_BINS)
    pooled = np.stack([h[:, :, s:e].mean(axis=2) for (s, e) in bins], axis=2)  # (B,32,8)
    flat = pooled.reshape(B, -1)
    st["bins"] = bins
    h = flat
    for name in TD_MLP:
        w_eff, sigma, u, v = sn_effective_weight(p, name, training)
        st["sn"][name] = (sigma, u, v)
        st["mlp_in"].append(h)
        h = leaky(h @ w_eff.T + p[name + ".bias"])
        st["mlp_out"].append(h)
        feats.append(h)
    if features_only:
        return feats, st
    w_eff, sigma, u, v = sn_effective_weight(p, TD_OUT, training)
    st["sn"][TD_OUT] = (sigma, u, v)
    st["out_in"] = h
    score = h @ w_eff.T + p[TD_OUT + ".bias"]
    return score, feats, st


def tdisc_bwd(p: Dict[str, Array], cfg: ModelCfg, st, dscore: Optional[Array], dfeats: Optional[List[Array]]):
    """Backward of one call.  Returns (grads wrt *_orig / bias, dx (B,T,3))."""
    grads: Dict[str, Array] = {}

    def eff(name):
        sigma, u, v = st["sn"][name]
        return p[name + ".weight_orig"] / sigma

    def put_w(name, g_eff):
        sigma, u, v = st["sn"][name]
        grads[name + ".weight_orig"] = sn_weight_grad(p[name + ".weight_orig"], g_eff, sigma, u, v)

    B = st["conv_in"][0].shape[0]
    if dscore is not None:
        h = st["out_in"]
        put_w(TD_OUT, dscore.T @ h)
        grads[TD_OUT + ".bias"] = dscore.sum(0)
        dh = dscore @ eff(TD_OUT)
    else:
        dh = np.zeros_like(st["mlp_out"][-1])
    for n in (1, 0):
        name = TD_MLP[n]
        if dfeats is not None:
            dh = dh + dfeats[3 + n]
        dpre = leaky_bwd(st["mlp_out"][n], dh)
        put_w(name, dpre.T @ st["mlp_in"][n])
        grads[name + ".bias"] = dpre.sum(0)
        dh = dpre @ eff(name)
    # un-pool
    c3 = st["conv_out"][2]
    dpool = dh.reshape(B, c3.shape[1], POOL_BINS)
    dconv = np.zeros_like(c3)
    for bi, (s, e) in enumerate(st["bins"]):
        dconv[:, :, s:e] += dpool[:, :, bi:bi + 1] / (e - s)
    dh = dconv
    for n in (2, 1, 0):
        name = TD_CONV[n]
        if dfeats is not None:
            dh = dh + dfeats[n].reshape(dh.shape)
        dpre = leaky_bwd(st["conv_out"][n], dh)
        dx, dw, db = conv1d_bwd(st["conv_in"][n], eff(name), dpre)
        put_w(name, dw)
        grads[name + ".bias"] = db
        dh = dx
    return grads, np.transpose(dh, (0, 2, 1))


# ----------------------------------------------------------------------------------------------
# MLP Discriminator (src/gan/models.py:182-243)
# ----------------------------------------------------------------------------------------------
def mdisc_layer_names(cfg: ModelCfg):
    return [f"layers.{i}" for i in range(len(cfg.disc_hidden_dims))]


def mdisc_fwd(p: Dict[str, Array], cfg: ModelCfg, x: Array, training: bool = True, features_only: bool = False):
    B = x.shape[0]
    h = x.reshape(B, -1)
    st = {"sn": {}, "ins": [], "outs": []}
    feats = []
    for name in mdisc_layer_names(cfg):
        w_eff, sigma, u, v = sn_effective_weight(p, name, training)
        st["sn"][name] = (sigma, u, v)
        st["ins"].append(h)
        h = leaky(h @ w_eff.T + p[name + ".bias"])
        st["outs"].append(h)
        feats.append(h)
    if features_only:
        return feats, st
    w_eff, sigma, u, v = sn_effective_weight(p, "output_layer", training)
    st["sn"]["output_layer"] = (sigma, u, v)
    st["out_in"] = h
    return h @ w_eff.T + p["output_layer.bias"], feats, st


def mdisc_bwd(p: Dict[str, Array], cfg: ModelCfg, st, dscore: Optional[Array], dfeats: Optional[List[Array]]):
    grads: Dict[str, Array] = {}
    names = mdisc_layer_names(cfg)

    def eff(name):
        return p[name + ".weight_orig"] / st["sn"][name][0]

    def put_w(name, g_eff):
        sigma, u, v = st["sn"][name]
        grads[name + ".weight_orig"] = sn_weight_grad(p[name + ".weight_orig"], g_eff, sigma, u, v)

    if dscore is not None:
        put_w("output_layer", dscore.T @ st["out_in"])
        grads["output_layer.bias"] = dscore.sum(0)
        dh = dscore @ eff("output_layer")
    else:
        dh = np.zeros_like(st["outs"][-1])
    for n in range(len(names) - 1, -1, -1):
        if dfeats is not None:
            dh = dh + dfeats[n]
        dpre = leaky_bwd(st["outs"][n], dh)
        put_w(names[n], dpre.T @ st["ins"][n])
        grads[names[n] + ".bias"] = dpre.sum(0)
        dh = dpre @ eff(names[n])
    B = dh.shape[0]
    return grads, dh.reshape(B, cfg.seq_length, cfg.input_dim)


def disc_fwd(p, cfg: ModelCfg, x, training=True, features_only=False):
    fn = tdisc_fwd if cfg.use_temporal_disc else mdisc_fwd  # src/gan/trainer.py:49
    return fn(p, cfg, x, training, features_only)


def disc_bwd(p, cfg: ModelCfg, st, dscore, dfeats):
    fn = tdisc_bwd if cfg.use_temporal_disc else mdisc_bwd
    return fn(p, cfg, st, dscore, dfeats)


# ----------------------------------------------------------------------------------------------
# Losses (src/gan/losses.py)
# ----------------------------------------------------------------------------------------------
def wasserstein_d(real_scores: Array, fake_scores: Array) -> float:
    return float(fake_scores.mean() - real_scores.mean())  # losses.py:43


def wasserstein_g(fake_scores: Array) -> float:
    return float(-fake_scores.mean())  # losses.py:58


def feature_matching(real_feats: List[Array], fake_feats: List[Array]):
    """losses.py:86-93.  Returns (loss, [dL/dfake_k]).  NB the double normalisation: l1 mean then / n_k."""
    K = len(real_feats)
    loss = 0.0
    grads = []
    for r, f in zip(real_feats, fake_feats):
        n_k = r.size / r.shape[0]
        diff = f - r
        loss += float(np.abs(diff).mean()) / n_k
        grads.append(np.sign(diff) / (diff.size * n_k * K))
    return loss / K, grads


def l1_mean(a: Array, b: Array):
    """F.l1_loss(a, b) (losses.py:120,147).  Returns (loss, dL/da)."""
    diff = a - b
    return float(np.abs(diff).mean()), np.sign(diff) / diff.size


def kl_divergence(mu: Array, log_var: Array):
    """losses.py:174-175.  Returns (loss, dmu, dlog_var)."""
    B = mu.shape[0]
    kld = -0.5 * np.sum(1.0 + log_var - mu * mu - np.exp(log_var), axis=1)
    return float(kld.mean()), mu / B, 0.5 * (np.exp(log_var) - 1.0) / B


# ----------------------------------------------------------------------------------------------
# clip_grad_norm_ + Adam (src/shared/utils.py:87-88; torch.optim.Adam, src/gan/trainer.py:60-79)
# ----------------------------------------------------------------------------------------------
def clip_grad_norm(grads: Dict[str, Array], max_norm: float) -> float:
    total = math.sqrt(sum(float((g.astype(np.float64) ** 2).sum()) for g in grads.values()))
    coef = min(max_norm / (total + 1e-6), 1.0)
    for k in grads:
        grads[k] = grads[k] * np.asarray(coef, grads[k].dtype)
    return total


def new_adam_state(params: Dict[str, Array], names: List[str]):
    return {
        "step": 0,
        "m": {k: np.zeros_like(params[k]) for k in names},
        "v": {k: np.zeros_like(params[k]) for k in names},
    }


def adam_step(params: Dict[str, Array], grads: Dict[str, Array], st, lr: float, tc: TrainCfg):
    b1, b2 = tc.betas
    st["step"] += 1
    t = st["step"]
    bc1 = 1.0 - b1 ** t
    bc2_sqrt = math.sqrt(1.0 - b2 ** t)
    for k, g in grads.items():
        m = st["m"][k]
        v = st["v"][k]
        m += (g - m) * (1.0 - b1)
        v *= b2
        v += (1.0 - b2) * g * g
        denom = np.sqrt(v) / bc2_sqrt + tc.adam_eps
        params[k] = params[k] - (lr / bc1) * (m / denom)


def trainable_names(p: Dict[str, Array]) -> List[str]:
    return [k for k in p if not (k.endswith("weight_u") or k.endswith("weight_v"))]


# ----------------------------------------------------------------------------------------------
# One full training batch, as written (src/shared/utils.py:62-139 + src/gan/trainer.py:84-193)
# ----------------------------------------------------------------------------------------------
@dataclass
class GanState:
    G: Dict[str, Array]
    E: Dict[str, Array]
    D1: Dict[str, Array]
    D2: Dict[str, Array]
    opt: Dict[str, dict] = field(default_factory=dict)

    def init_opt(self):
        for n in ("G", "E", "D1", "D2"):
            p = getattr(self, n)
            self.opt[n] = new_adam_state(p, trainable_names(p))


def n_noise_draws(tc: TrainCfg) -> int:
    return 2 * tc.n_critic + 3


def train_batch(
    s: GanState,
    cfg: ModelCfg,
    tc: TrainCfg,
    real: Array,
    proto: Array,
    noise: List[Array],
    max_norm: float = 1.0,
    lrs: Optional[Dict[str, float]] = None,
    record: Optional[dict] = None,
):
    """One DataLoader batch of train_epoch_with_grad_clip (utils.py:62-139).

    ``noise`` holds the 2*n_critic+3 (B,Z) normal draws in the reference's consumption order:
    per critic iteration z_rand (utils.py:71) then the encoder's eps (utils.py:93 -> models.py:85);
    then cycle-1 z (trainer.py:105), eps of the no-grad recovery pass (trainer.py:118), eps of
    cycle 2 (trainer.py:161).  Mutates ``s``; returns the loss dict (all 11 logged scalars).
    If ``record`` is a dict it receives the clipped-before gradients of each optimiser step.
    """
    lrs = lrs or {k: tc.learning_rate for k in ("G", "E", "D1", "D2")}
    losses: Dict[str, float] = {}
    B = real.shape[0]
    ni = 0
    for it in range(tc.n_critic):
        for which, Dp in (("D1", s.D1), ("D2", s.D2)):
            if which == "D1":
                z = noise[ni]
                ni += 1
            else:
                z, _, _, _ = encoder_fwd(s.E, cfg, real, noise[ni])
                ni += 1
            fake, _ = generator_fwd(s.G, cfg, proto, z, want_stash=False)
            rs, _, st_r = disc_fwd(Dp, cfg, real, True)  # utils.py:77 / :98
            fs, _, st_f = disc_fwd(Dp, cfg, fake, True)  # utils.py:78 / :99
            losses["d1_loss" if which == "D1" else "d2_loss"] = wasserstein_d(rs, fs)
            g_r, _ = disc_bwd(Dp, cfg, st_r, np.full_like(rs, -1.0 / B), None)
            g_f, _ = disc_bwd(Dp, cfg, st_f, np.full_like(fs, 1.0 / B), None)
            grads = {k: g_r[k] + g_f[k] for k in g_r}
            if record is not None:
                record[f"{which}_grads_{it}"] = {k: v.copy() for k, v in grads.items()}
            clip_grad_norm(grads, max_norm)
            adam_step(Dp, grads, s.opt[which], lrs[which], tc)

    # ---- generator / encoder step (utils.py:112-135) ----
    # cycle 1 (trainer.py:102-131)
    z1 = noise[ni]
    ni += 1
    fake1, gst1 = generator_fwd(s.G, cfg, proto, z1)
    sc1, _, st_s1 = disc_fwd(s.D1, cfg, fake1, True)
    ff1, st_ff1 = disc_fwd(s.D1, cfg, fake1, True, features_only=True)
    rf1, _ = disc_fwd(s.D1, cfg, real, True, features_only=True)
    z_rec, _, _, _ = encoder_fwd(s.E, cfg, fake1, noise[ni])  # no-grad pass (trainer.py:116-119)
    ni += 1
    l_wgan1 = wasserstein_g(sc1)
    l_feat1, dff1 = feature_matching(rf1, ff1)
    l_lat, _ = l1_mean(z_rec, z1)
    losses["cycle1_wgan"] = l_wgan1
    losses["cycle1_feat"] = l_feat1
    losses["cycle1_lat"] = l_lat
    losses["cycle1_total"] = l_wgan1 + tc.lambda_feat * l_feat1 + tc.lambda_lat * l_lat
    _, dx_a = disc_bwd(s.D1, cfg, st_s1, np.full_like(sc1, -1.0 / B), None)
    _, dx_b = disc_bwd(s.D1, cfg, st_ff1, None, [tc.lambda_feat * g for g in dff1])
    gG1, _ = generator_bwd(s.G, cfg, gst1, dx_a + dx_b)

    # cycle 2 (trainer.py:160-183)
    z2, mu, log_var, est = encoder_fwd(s.E, cfg, real, noise[ni])
    ni += 1
    fake2, gst2 = generator_fwd(s.G, cfg, proto, z2)
    sc2, _, st_s2 = disc_fwd(s.D2, cfg, fake2, True)
    ff2, st_ff2 = disc_fwd(s.D2, cfg, fake2, True, features_only=True)
    rf2, _ = disc_fwd(s.D2, cfg, real, True, features_only=True)
    l_wgan2 = wasserstein_g(sc2)
    l_feat2, dff2 = feature_matching(rf2, ff2)
    l_rec, drec = l1_mean(fake2, real)
    l_kld, dmu, dlv = kl_divergence(mu, log_var)
    losses["cycle2_wgan"] = l_wgan2
    losses["cycle2_feat"] = l_feat2
    losses["cycle2_rec"] = l_rec
    losses["cycle2_kld"] = l_kld
    losses["cycle2_total"] = l_wgan2 + tc.lambda_feat * l_feat2 + tc.lambda_rec * l_rec + tc.lambda_kld * l_kld
    _, dx_a = disc_bwd(s.D2, cfg, st_s2, np.full_like(sc2, -1.0 / B), None)
    _, dx_b = disc_bwd(s.D2, cfg, st_ff2, None, [tc.lambda_feat * g for g in dff2])
    gG2, dz2 = generator_bwd(s.G, cfg, gst2, dx_a + dx_b + tc.lambda_rec * drec)
    gE = encoder_bwd(s.E, cfg, est, dz2, tc.lambda_kld * dmu, tc.lambda_kld * dlv)
    gG = {k: gG1[k] + gG2[k] for k in gG1}
    if record is not None:
        record["G_grads"] = {k: v.copy() for k, v in gG.items()}
        record["E_grads"] = {k: v.copy() for k, v in gE.items()}
        record["fake_cycle1"] = fake1
        record["fake_cycle2"] = fake2
    clip_grad_norm(gG, max_norm)  # utils.py:132
    clip_grad_norm(gE, max_norm)  # utils.py:133
    adam_step(s.G, gG, s.opt["G"], lrs["G"], tc)
    adam_step(s.E, gE, s.opt["E"], lrs["E"], tc)
    assert ni == n_noise_draws(tc)
    return losses


def sample(G: Dict[str, Array], cfg: ModelCfg, proto: Array, z: Array) -> Array:
    """Sampling hot path (eval_gan.py:131-135): generator(protos, z) in eval / no-grad."""
    return generator_fwd(G, cfg, proto, z, want_stash=False)[0]
