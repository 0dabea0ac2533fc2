"""Helpers shared by the tests: load the reference-generated fixtures (tests/golden/*.npz, produced by
oracle/make_golden.py from the unmodified reference) and compare tensors against them."""
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ("tiny_temporal", "tiny_mlp_time", "default")
MODS = ("G", "E", "D1", "D2")
LOSS_KEYS = ("d1_loss", "d2_loss", "cycle1_wgan", "cycle1_feat", "cycle1_lat", "cycle1_total", "cycle2_wgan",
             "cycle2_feat", "cycle2_rec", "cycle2_kld", "cycle2_total")


class Golden:
    def __init__(self, name):
        self.name = name
        self.z = np.load(os.path.join(GOLDEN_DIR, f"step_{name}.npz"), allow_pickle=False)
        self.keys = set(self.z.files)

    def cfg_kwargs(self):
        z = self.z
        return dict(seq_length=int(z["cfg/seq_length"]), latent_dim=int(z["cfg/latent_dim"]),
                    gen_hidden_dim=int(z["cfg/gen_hidden_dim"]), gen_num_layers=int(z["cfg/gen_num_layers"]),
                    enc_hidden_dims=tuple(int(v) for v in z["cfg/enc_hidden_dims"]),
                    disc_hidden_dims=tuple(int(v) for v in z["cfg/disc_hidden_dims"]),
                    use_temporal_disc=bool(int(z["cfg/use_temporal_disc"])),
                    prototype_has_time=bool(int(z["cfg/prototype_has_time"])))

    def init_state(self, mod, dtype=np.float64):
        order = [str(k) for k in self.z[f"order/{mod}"]]
        return {k: self.z[f"init/{mod}/{k}"].astype(dtype) for k in order}

    def param_order(self, mod):
        return [str(k) for k in self.z[f"porder/{mod}"]]

    def inputs(self, dtype=np.float64):
        return (self.z["real"].astype(dtype), self.z["proto"].astype(dtype),
                [n.astype(dtype) for n in self.z["noise"]])

    def loss(self, k):
        return float(self.z[f"loss/{k}"])

    def has(self, kind, grp, key):
        return f"full/{kind}/{grp}/{key}" in self.keys or f"sum/{kind}/{grp}/{key}" in self.keys

    def check(self, kind, grp, key, value, tol, what=""):
        """Compare ``value`` with the stored tensor (full array) or its summary; returns the error measured."""
        v = np.asarray(value, np.float64)
        fk, sk = f"full/{kind}/{grp}/{key}", f"sum/{kind}/{grp}/{key}"
        if fk in self.keys:
            ref = self.z[fk]
            assert ref.shape == v.shape, (what, kind, grp, key, ref.shape, v.shape)
            err = rel_l2(v, ref)
        else:
            ref = self.z[sk]
            mine = summarise(v)
            scale = max(ref[2], 1e-30)  # l2 norm of the reference tensor
            n = v.size
            # sampled entries: error relative to the tensor's RMS; sums: relative to l2*sqrt(n) (Cauchy-Schwarz bound)
            rms = scale / np.sqrt(n)
            e_samples = np.abs(mine[3:] - ref[3:]).max() / max(rms, 1e-30) / np.sqrt(n) * np.sqrt(min(n, 48))
            e_l2 = abs(mine[2] - ref[2]) / scale
            err = max(e_l2, e_samples)
        assert err <= tol, f"{what} {kind}/{grp}/{key}: error {err:.3e} > tol {tol:.1e}"
        return err


def rel_l2(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    d = np.sqrt(((a - b) ** 2).sum())
    n = np.sqrt((b ** 2).sum())
    return d / n if n > 0 else d


def max_abs_rel(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    m = np.abs(b).max()
    return np.abs(a - b).max() / (m if m > 0 else 1.0)


def summarise(a):
    a = np.asarray(a, np.float64)
    flat = a.ravel()
    idx = np.linspace(0, flat.size - 1, min(48, flat.size)).astype(np.int64)
    return np.concatenate([[a.sum(), np.abs(a).sum(), np.sqrt((a * a).sum())], flat[idx]])


def oracle_cfg(g):
    from oracle import wgg_oracle as O
    return O.ModelCfg(**g.cfg_kwargs())


class CycleGolden:
    """tests/golden/cycles_<case>.npz: direct calls of the reference's train_generator_step_cycle1/2
    (oracle/make_golden_cycles.py)."""

    def __init__(self, name):
        self.name = name
        self.z = np.load(os.path.join(GOLDEN_DIR, f"cycles_{name}.npz"), allow_pickle=False)

    cfg_kwargs = Golden.cfg_kwargs

    def init_state(self, mod, dtype=np.float64):
        return {str(k): self.z[f"init/{mod}/{k}"].astype(dtype) for k in self.z[f"order/{mod}"]}

    def inputs(self):
        n = self.z["noise"]
        return self.z["real"], self.z["proto"], n[0], n[1], n[2]

    def losses(self, cyc):
        pre = f"c{cyc}/dict/"
        return {k[len(pre):]: float(self.z[k]) for k in self.z.files if k.startswith(pre)}

    def grads(self, cyc, mod):
        pre = f"c{cyc}/grad/{mod}/"
        return {k[len(pre):]: self.z[k] for k in self.z.files if k.startswith(pre)}

    def uv(self, cyc):
        pre = f"c{cyc}/uv/"
        return {k[len(pre):]: self.z[k] for k in self.z.files if k.startswith(pre)}
