"""Times the no-grad generator forward (critic-phase call of 10 x 4096 gestures, and sampling calls) for an A/B of a
kernel knob given through the environment (e.g. WGG_FWD_NCH=2 vs 3)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import wgg_b200 as wgg
from wgg_b200 import _lib
dev = torch.device("cuda:0")
wgg.set_math_mode("tf32"); wgg.seed_everything(42)
G = wgg.Generator().to(dev).eval()
knobs = {k: v for k, v in os.environ.items() if k.startswith("WGG_")}
for B in (9472, 18944, 40960):
    proto = torch.rand(B, 128, 3, device=dev) * 2 - 1
    z = torch.randn(B, 32, device=dev)
    with torch.no_grad():
        for _ in range(3): y = G(proto, z)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): y = G(proto, z)
        e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{knobs} B={B}: {ms:.3f} ms/call, {B/ms*1e3/1e6:.2f} M samples/s, checksum {y.double().sum().item():.6f} "
          f"async_err {_lib.async_error(dev)}", flush=True)
