"""BASELINE configs[4]: generator-only sampling of 1 M gestures (eval / no-grad), output kept in HBM, then one D2H."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import wgg_b200 as wgg
dev = torch.device("cuda:0")
wgg.set_math_mode("tf32"); wgg.seed_everything(42)
gen = wgg.Generator(wgg.ModelConfig()).to(dev).eval()
N, BS = 1_000_000, 74 * 128
protos = (torch.rand(N, 128, 3, generator=torch.Generator().manual_seed(0)) * 2 - 1)
protos_d = protos.to(dev)
out = torch.empty(N, 128, 3, device=dev)
def run():
    with torch.no_grad():
        for lo in range(0, N, BS):
            hi = min(lo + BS, N)
            out[lo:hi] = gen(protos_d[lo:hi], torch.randn(hi - lo, 32, device=dev))
run(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
t0 = time.perf_counter(); host = out.cpu(); d2h = time.perf_counter() - t0
print(json.dumps({"config": "1M generator samples, 1 x B200, batch 9472 per call, tf32", "ms": ms, "samples_per_s": N / ms * 1e3,
                  "d2h_seconds_1.5GB": d2h, "finite": bool(torch.isfinite(host).all()), "range": [float(host.min()), float(host.max())]}))
