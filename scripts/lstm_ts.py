"""Debug: per-step clock stamps of lstm_tc_fwd_kernel CTA (0,0) during a no-grad generator forward."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import wgg_b200 as wgg
from wgg_b200 import _lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 40960
dev = torch.device("cuda:0")
wgg.set_math_mode("tf32"); wgg.seed_everything(42)
tr = wgg.WordGestureGANTrainer(wgg.ModelConfig(), wgg.TrainingConfig(), dev)
tr.generator.eval()
proto = torch.rand(B, 128, 3, device=dev) * 2 - 1
z = torch.randn(B, 32, device=dev)
lib = _lib.lib() if callable(getattr(_lib, "lib", None)) else _lib._LIB
lib.wgg_debug_lstm_ts.argtypes = [ctypes.c_int, ctypes.c_void_p]
with torch.no_grad():
    tr.generator(proto, z); torch.cuda.synchronize()
    lib.wgg_debug_lstm_ts(1, None)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); tr.generator(proto, z); e1.record(); torch.cuda.synchronize()
    print("generator forward ms", e0.elapsed_time(e1))
out = np.zeros(1024, dtype=np.int64)
lib.wgg_debug_lstm_ts(0, out.ctypes.data_as(ctypes.c_void_p))
ts = out.reshape(128, 8)
base = ts[0, 0]
names = ["mma:acc_empty", "mma:x_issued", "mma:h_ready", "mma:committed", "epi:acc_full", "epi:tmem_read", "epi:computed", "epi:h_arrived"]
print("step " + " ".join(f"{n:>14s}" for n in names))
for s in list(range(0, 6)) + list(range(60, 64)) + list(range(124, 128)):
    print(f"{s:4d} " + " ".join(f"{int(v - base):14d}" for v in ts[s]))
d = np.diff(ts[:, 4])
print("epi acc_full period: mean", d[2:].mean(), "min", d[2:].min(), "max", d[2:].max())
for a, b, lab in [(4, 5, "tmem read"), (5, 6, "gate math+stores"), (6, 7, "fence+arrive")]:
    print(lab, (ts[2:, b] - ts[2:, a]).mean())
print("h_arrived(warp2) -> mma h_ready(next step)", (ts[3:, 2] - ts[2:-1, 7]).mean())
print("mma h_ready -> committed", (ts[2:, 3] - ts[2:, 2]).mean())
print("mma committed -> epi acc_full", (ts[2:, 4] - ts[2:, 3]).mean())
