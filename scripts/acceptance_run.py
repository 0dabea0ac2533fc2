"""Acceptance run (north star: "after a short training run the reference's own eval_gan.py metrics ... within noise").

On the GPU box: train (a) this implementation (math mode tf32, CUDA-graph step, device-resident loader) and (b) the
UNMODIFIED reference on cuda (its own trainer + train_epoch_with_grad_clip, cuDNN / cuBLAS; oracle/ref_runner.py) on the
same realistic fixture (tests/golden/realistic_gestures.npz, made by oracle/make_realistic_fixture.py with the
reference's keyboard / minimum-jerk code) for the same short schedule (batch 512, cosine LR, TRAIN_SCRIPT's recipe:
train_gan.py:95-100,150), several seeds each; generate gestures for the held-out prototypes (eval_gan.py:123-135) and
feed every model's output to the reference's OWN evaluate_all_metrics(skip_dtw=True) (src/gan/evaluation.py:297) with
one shared FID auto-encoder.  Free-running training cannot be compared number by number, so the check
is statistical: for every metric, |mean_ours - mean_reference| against the seed-to-seed spread.

  python scripts/acceptance_run.py --epochs 100 --seeds 0,1,2 --out gpurun_out/r02_acceptance.json
"""
import argparse
import json
import os
import sys
import time
from pathlib import Path

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch
from torch.optim.lr_scheduler import CosineAnnealingLR

METRICS = ("l2_wasserstein", "jerk_fake", "velocity_corr", "acceleration_corr", "speed_profile_corr", "time_delta_corr",
           "fid", "precision", "recall")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--epochs", type=int, default=100)
    ap.add_argument("--seeds", default="0,1,2")
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--math", default="tf32")
    ap.add_argument("--out", default="gpurun_out/r02_acceptance.json")
    args = ap.parse_args()
    seeds = [int(s) for s in args.seeds.split(",")]
    dev = torch.device("cuda:0")
    fx = np.load(os.path.join(ROOT, "tests", "golden", "realistic_gestures.npz"))
    n_train = (fx["train_gesture"].shape[0] // args.batch) * args.batch     # full batches only: same schedule both sides
    train_g = torch.from_numpy(fx["train_gesture"][:n_train].astype(np.float32))
    train_p = torch.from_numpy(fx["train_prototype"][:n_train].astype(np.float32))
    test_g = fx["test_gesture"].astype(np.float32)
    test_p = torch.from_numpy(fx["test_prototype"].astype(np.float32))
    n_test = test_g.shape[0]

    from oracle.ref_runner import ReferenceRunner
    from oracle.ref_loader import load_reference
    ref = load_reference()
    os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
    # evaluate_all_metrics caches its FID auto-encoder under /data (Modal volume): redirect, in this process only
    cache_dir = Path(os.path.dirname(os.path.abspath(args.out)))
    ref.evaluation._get_ae_cache_path = lambda train_data, eval_config: cache_dir / "r02_fid_ae_cache.pt"
    if (cache_dir / "r02_fid_ae_cache.pt").exists():
        (cache_dir / "r02_fid_ae_cache.pt").unlink()
    cached = {}

    def evaluate(fake):
        torch.manual_seed(1234)  # the FID auto-encoder is trained once (first call), then shared through cached_real
        res = ref.evaluation.evaluate_all_metrics(test_g, fake, train_gestures=train_g.numpy(), device="cuda", skip_dtw=True,
                                                  cached_real=cached.get("c"))
        cached["c"] = res.pop("_cached_real")
        return {k: float(v) for k, v in res.items()}

    def z_for(seed):
        return torch.randn(n_test, 32, generator=torch.Generator().manual_seed(10_000 + seed))

    runs = {"ours": [], "reference_cuda": [], "untrained": []}
    timing = {}

    # ---- (b) the reference on cuda ----
    for seed in seeds:
        rr = ReferenceRunner(dev, seed=42 + seed, batch_size=args.batch)
        tr = rr.trainer
        if seed == seeds[0]:
            runs["untrained"].append(evaluate(rr.sample(test_p, z_for(seed)).cpu().numpy()))
        scheds = [CosineAnnealingLR(o, T_max=args.epochs, eta_min=1e-5)
                  for o in (tr.optimizer_G, tr.optimizer_E, tr.optimizer_D1, tr.optimizer_D2)]
        gd, pd = train_g.to(dev), train_p.to(dev)
        gen = torch.Generator(device=dev).manual_seed(42 + seed)
        torch.cuda.synchronize()
        t0 = time.time()
        last = None
        for ep in range(args.epochs):
            perm = torch.randperm(n_train, device=dev, generator=gen)
            batches = [{"gesture": gd[perm[i:i + args.batch]], "prototype": pd[perm[i:i + args.batch]]}
                       for i in range(0, n_train, args.batch)]
            last = rr.train_batches(batches)
            for s in scheds:
                s.step()
        torch.cuda.synchronize()
        timing.setdefault("reference_cuda_train_s", []).append(time.time() - t0)
        m = evaluate(rr.sample(test_p, z_for(seed)).cpu().numpy())
        m["final_losses"] = {k: float(v) for k, v in last.items()}
        runs["reference_cuda"].append(m)
        print("reference_cuda seed", seed, {k: round(m[k], 4) for k in METRICS}, flush=True)
        del rr, tr

    # ---- (a) this implementation ----
    import wgg_b200 as wgg
    import tempfile
    for seed in seeds:
        with tempfile.TemporaryDirectory() as ck:
            torch.cuda.synchronize()
            t0 = time.time()
            tc = wgg.TrainingConfig(batch_size=args.batch, num_epochs=args.epochs)
            hist = wgg.run_training(train_g, train_p, args.epochs, ck, resume=False, training_config=tc, seed=42 + seed,
                                    device=dev, use_cuda_graph=True, verbose=False, math_mode=args.math,
                                    checkpoint_every=10 ** 9)
            torch.cuda.synchronize()
            timing.setdefault("ours_train_s", []).append(time.time() - t0)
            ckpt = torch.load(os.path.join(ck, "latest.pt"), map_location=dev)
        G = wgg.Generator().to(dev)
        G.load_state_dict(ckpt["generator"])
        G.eval()
        with torch.no_grad():
            fake = G(test_p.to(dev), z_for(seed).to(dev)).cpu().numpy()
        m = evaluate(fake)
        m["final_losses"] = {k: float(hist[-1][k]) for k in ("d1_loss", "d2_loss", "cycle1_total", "cycle2_total")}
        if seed == seeds[0]:
            # the same arrays through this package's GPU metric kernels, next to the reference's
            ae = cached["c"]["autoencoder"]
            gm = wgg.eval_metrics.evaluate_all_metrics(test_g, fake, device=dev, autoencoder=ae)
            m["gpu_metrics"] = gm
            m["gpu_metrics_abs_diff"] = {k: abs(gm[k] - m[k]) for k in METRICS if k in gm}
        runs["ours"].append(m)
        print("ours seed", seed, {k: round(m[k], 4) for k in METRICS}, flush=True)

    table = {}
    for k in METRICS:
        a = np.array([r[k] for r in runs["ours"]])
        b = np.array([r[k] for r in runs["reference_cuda"]])
        spread = float(np.sqrt((a.var(ddof=1) + b.var(ddof=1)) / 2)) if len(a) > 1 else float("nan")
        table[k] = {"ours_mean": float(a.mean()), "ours_std": float(a.std(ddof=1)) if len(a) > 1 else None,
                    "reference_mean": float(b.mean()), "reference_std": float(b.std(ddof=1)) if len(b) > 1 else None,
                    "untrained": runs["untrained"][0][k], "abs_diff_of_means": float(abs(a.mean() - b.mean())),
                    "pooled_seed_std": spread,
                    # the two means differ by less than 3 standard errors of their difference, or by less than 2 % of
                    # the distance the metric travels from the untrained model to the trained ones
                    "within_noise": bool(abs(a.mean() - b.mean()) <= max(3 * spread * np.sqrt(2 / len(a)),
                                                                          0.02 * abs(runs["untrained"][0][k] - b.mean())) + 1e-12)
                    if len(a) > 1 else None}
    out = {"what": __doc__.split("\n")[0], "epochs": args.epochs, "batch": args.batch, "seeds": seeds, "math_mode": args.math,
           "train_gestures": n_train, "test_gestures": n_test, "steps_per_run": args.epochs * (n_train // args.batch),
           "real_jerk": runs["ours"][0].get("jerk_real"), "timing_s": timing, "table": table, "runs": runs}
    json.dump(out, open(args.out, "w"), indent=1)
    print(json.dumps({"table": table, "timing_s": timing}, indent=1))


if __name__ == "__main__":
    main()
