"""Debug: clock stamps of lstm_tc_dw_kernel CTA (0,0) during one generator backward."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import wgg_b200 as wgg
from wgg_b200 import _lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
dev = torch.device("cuda:0")
wgg.set_math_mode("tf32"); wgg.seed_everything(42)
tr = wgg.WordGestureGANTrainer(wgg.ModelConfig(), wgg.TrainingConfig(), dev)
tr.generator.train()
proto = torch.rand(B, 128, 3, device=dev) * 2 - 1
z = torch.randn(B, 32, device=dev, requires_grad=True)
lib = _lib.lib()
lib.wgg_debug_lstm_ts.argtypes = [ctypes.c_int, ctypes.c_void_p]
out = tr.generator(proto, z); out.sum().backward(); torch.cuda.synchronize()
lib.wgg_debug_lstm_ts(1, None)
out = tr.generator(proto, z); out.sum().backward(); torch.cuda.synchronize()
buf = np.zeros(1024, dtype=np.int64)
lib.wgg_debug_lstm_ts(0, buf.ctypes.data_as(ctypes.c_void_p))
ts = buf.reshape(128, 8)
names = ["mma:full", "mma:issued", "tr:raw_full", "tr:op_empty", "tr:stored", "tr:arrived", "ld:raw_empty", "ld:issued"]
base = ts[:, [0, 2, 6]].min()
print("n    " + " ".join(f"{x:>12s}" for x in names))
for n in list(range(8, 24)):
    print(f"{n:4d} " + " ".join(f"{int(v - base):12d}" for v in ts[n]))
print("period (mma:full)", np.diff(ts[8:120, 0]).mean())
for a, b, lab in [(2, 3, "tr wait op_empty"), (3, 4, "tr transpose+store"), (4, 5, "tr fence+arrive"), (0, 1, "mma issue"), (6, 7, "loader issue")]:
    print(lab, (ts[8:120, b] - ts[8:120, a]).mean())
print("tr arrived -> mma full", (ts[8:120, 0] - ts[8:120, 5]).mean())
print("mma issued(n) -> tr op_empty(n+2)", (ts[10:120, 3] - ts[8:118, 1]).mean())
print("ld issued(n) -> tr raw_full(n)", (ts[8:120, 2] - ts[8:120, 7]).mean())
