import sys, os, ctypes
os.environ["WGG_DEBUG_WGRAD_DUMP"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import wgg_b200 as wgg
from wgg_b200 import _lib
dev = torch.device("cuda:0")
lib = _lib.lib(); c = _lib.ctx(dev)
P = ctypes.c_void_p
lib.wgg_debug_conv_tc_wgrad.restype = ctypes.c_int
lib.wgg_debug_conv_tc_wgrad.argtypes = [P, P, P, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, P, P, P, P]
T = 128
def chunk(x):
    B, T_, C = x.shape
    C4 = (C + 3) // 4 * 4
    xp = torch.zeros(B, T_, C4, device=x.device); xp[:, :, :C] = x
    return xp.view(B, T_, C4 // 4, 4).permute(0, 2, 1, 3).contiguous()
def run(dpre, x, Cout, Cin, taps, pad):
    B = dpre.shape[0]
    G = torch.zeros((Cout, taps * Cin), device=dev); db = torch.zeros(Cout, device=dev)
    ws = torch.zeros((256 * 128 * 336,), device=dev)
    lib.wgg_debug_conv_tc_wgrad(c, chunk(dpre).data_ptr(), chunk(x).data_ptr(), B, Cout, Cin, taps, pad, G.data_ptr(), db.data_ptr(), ws.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return ws
Cout, Cin, taps, pad = 64, 64, 5, 2
ncols = taps * 64 + 8
import itertools
for mask, (t0, c0, t1, c1) in itertools.product([1 | 4, 4 | 2 | 8, 1, 2 | 8, 2], [(40, 5, 41, 2), (42, 37, 42, 50)]):
    os.environ["WGG_DEBUG_WGRAD_MASK"] = str(mask)
    dpre = torch.zeros(1, T, Cout, device=dev); x = torch.zeros(1, T, Cin, device=dev)
    dpre[0, t0, c0] = 1.0; x[0, t1, c1] = 2.0
    ws = run(dpre, x, Cout, Cin, taps, pad)
    part = ws[:128 * ncols].view(128, ncols)
    nz = (part != 0).nonzero().tolist()
    tap = t1 - t0 + pad
    print("MASK", mask, end=" ")
    print(f"dpre[t={t0}][co={c0}]=1  x[t={t1}][ci={c1}]=2  expect row(co)={c0} -> lane {(c0%16)+32*(c0//16)}, col={tap*64+c1} (tap {tap}) val 2; bias cols {taps*64}.. val 1")
    print("   lanes/cols:", [(r, cc, part[r, cc].item()) for r, cc in nz[:14]], "count", len(nz))
