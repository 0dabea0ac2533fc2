"""Golden fixtures for DIRECT calls of the reference's generator-side cycle builders.

Run in the build container only (needs /root/reference):  python -m oracle.make_golden_cycles

``WordGestureGANTrainer.train_generator_step_cycle1`` / ``..._cycle2`` (src/gan/trainer.py:84-140, :142-193) of the
UNMODIFIED reference are called once each, in float64 on CPU, from the seed-42 initial state (rounded to fp32
representable values), with ``torch.randn`` / ``torch.randn_like`` patched in this process so the three normal
draws (cycle-1 z, the no-grad recovery eps, cycle-2 eps) are explicit inputs.  After each call the returned total
loss is back-propagated on its own and the generator / encoder gradients are recorded, together with the returned
fake gesture, the returned loss dict and the discriminator's spectral-norm buffers after its three calls.
Written to tests/golden/cycles_<case>.npz; tests/test_oracle_golden.py pins the CPU restatements to it and
tests/test_gpu_parity_tc.py compares the product's public methods with it.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle.make_golden import CASES, GOLDEN_DIR, sd_to_np  # noqa: E402
from oracle.ref_loader import load_reference  # noqa: E402

MODS = (("G", "generator"), ("E", "encoder"), ("D1", "discriminator_1"), ("D2", "discriminator_2"))


def run_case(ref, mcfg_kwargs, B, seed):
    torch.set_default_dtype(torch.float64)
    try:
        mcfg = ref.config.ModelConfig(**mcfg_kwargs)
        tcfg = ref.config.TrainingConfig()
        ref.utils.seed_everything(42)
        trainer = ref.trainer.WordGestureGANTrainer(mcfg, tcfg, device="cpu")
        with torch.no_grad():
            for _, m in MODS:
                mod = getattr(trainer, m)
                mod.train()
                for t in list(mod.parameters()) + list(mod.buffers()):
                    t.copy_(t.float().double())
        out = {}
        for n, m in MODS:
            sd = sd_to_np(getattr(trainer, m))
            out[f"order/{n}"] = np.array(list(sd.keys()))
            for k, v in sd.items():
                out[f"init/{n}/{k}"] = v.astype(np.float32)
        g = torch.Generator().manual_seed(100 + seed)
        T, Z = mcfg.seq_length, mcfg.latent_dim
        real = (torch.rand(B, T, 3, generator=g) * 2 - 1).float().double()
        proto = (torch.rand(B, T, 3, generator=g) * 2 - 1).float().double()
        noise = [torch.randn(B, Z, generator=g).float().double() for _ in range(3)]  # z, eps_recover, eps
        out.update(real=real.numpy(), proto=proto.numpy(), noise=np.stack([n.numpy() for n in noise]))
        for k in ("seq_length", "latent_dim", "gen_hidden_dim", "gen_num_layers"):
            out[f"cfg/{k}"] = getattr(mcfg, k)
        out["cfg/enc_hidden_dims"] = np.array(mcfg.enc_hidden_dims)
        out["cfg/disc_hidden_dims"] = np.array(mcfg.disc_hidden_dims)
        out["cfg/use_temporal_disc"] = int(mcfg.use_temporal_disc)
        out["cfg/prototype_has_time"] = int(mcfg.prototype_has_time)

        queue = [n.clone() for n in noise]
        orig_randn, orig_randn_like = torch.randn, torch.randn_like

        def fake_randn(*a, **k):
            return queue.pop(0)

        def fake_randn_like(t, *a, **k):
            o = queue.pop(0)
            assert o.shape == t.shape
            return o

        torch.randn, torch.randn_like = fake_randn, fake_randn_like
        try:
            for cyc, fn, disc in ((1, trainer.train_generator_step_cycle1, "discriminator_1"),
                                  (2, trainer.train_generator_step_cycle2, "discriminator_2")):
                trainer.optimizer_G.zero_grad()
                trainer.optimizer_E.zero_grad()
                fake, total, d = fn(proto, real)
                total.backward()
                out[f"c{cyc}/fake"] = fake.detach().numpy().copy()
                out[f"c{cyc}/total"] = np.float64(total.item())
                for k, v in d.items():
                    assert isinstance(v, float)
                    out[f"c{cyc}/dict/{k}"] = np.float64(v)
                for n, m in (("G", "generator"), ("E", "encoder")):
                    for k, p in getattr(trainer, m).named_parameters():
                        out[f"c{cyc}/grad/{n}/{k}"] = (np.zeros(tuple(p.shape)) if p.grad is None
                                                        else p.grad.detach().numpy().copy())
                for k, v in getattr(trainer, disc).state_dict().items():
                    if k.endswith("weight_u") or k.endswith("weight_v"):
                        out[f"c{cyc}/uv/{k}"] = v.detach().numpy().copy()
        finally:
            torch.randn, torch.randn_like = orig_randn, orig_randn_like
        assert not queue
        return out
    finally:
        torch.set_default_dtype(torch.float32)


def main():
    ref = load_reference()
    for name, (kw, B, seed) in CASES.items():
        out = run_case(ref, kw, B, seed)
        path = os.path.join(GOLDEN_DIR, f"cycles_{name}.npz")
        np.savez_compressed(path, **out)
        print(f"[golden] wrote {path} ({os.path.getsize(path) / 1e6:.2f} MB)")


if __name__ == "__main__":
    main()
