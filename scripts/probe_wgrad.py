"""Probe the tcgen05 conv weight-gradient kernel with delta inputs to decode operand / TMEM layouts."""
import sys, os, ctypes
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
def chunk(x):  # (B,T,C) -> [B][C/4][T][4], C padded to 4
    B, T_, C = x.shape
    C4 = (C + 3) // 4 * 4
    xp = torch.zeros(B, T_, C4, device=x.device); xp[:, :, :C] = x
    return xp.view(B, T_, C4 // 4, 4).permute(0, 2, 1, 3).contiguous()
def run(dpre, x, Cout, Cin, taps, pad):
    B = dpre.shape[0]
    G = torch.full((Cout, taps * Cin), -7.0, device=dev); db = torch.zeros(Cout, device=dev)
    ws = torch.full((256 * 64 * 336,), -3.0, device=dev)
    dpc, xc = chunk(dpre), chunk(x)  # keep the temporaries alive: the kernel reads them asynchronously
    rc = lib.wgg_debug_conv_tc_wgrad(c, dpc.data_ptr(), xc.data_ptr(), B, Cout, Cin, taps, pad, G.data_ptr(), db.data_ptr(), ws.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    print("rc", rc, "async_err", _lib.async_error(dev))
    return G, db, ws
for (Cout, Cin, taps, pad) in [(64, 64, 5, 2), (32, 64, 3, 1), (64, 3, 5, 2)]:
    print("=== config", Cout, Cin, taps, pad)
    B = 1
    dpre = torch.zeros(B, T, Cout, device=dev); x = torch.zeros(B, T, Cin, device=dev)
    t0, c0, c1 = 40, 5, 2
    dpre[0, t0, c0] = 1.0
    x[0, t0 + 1, c1] = 2.0      # tap = pad + 1 should see it
    G, db, ws = run(dpre, x, Cout, Cin, taps, pad)
    nz = (G != 0).nonzero().tolist()
    print("G nonzeros (expect [[%d, %d]] = 2.0):" % (c0, (pad + 1) * Cin + c1), [(r, cc, G[r, cc].item()) for r, cc in nz[:12]], "count", len(nz))
    print("db nonzeros (expect [%d]=1):" % c0, [(i, db[i].item()) for i in (db != 0).nonzero().flatten().tolist()[:8]])
    ncols = (taps * ((Cin + 3) // 4 * 4) if Cin > 4 else (taps + taps % 2) * 4) + 8
    part = ws[:64 * ncols].view(64, ncols)
    pnz = (part != 0).nonzero().tolist()
    print("partial nonzeros:", [(r, cc, part[r, cc].item()) for r, cc in pnz[:16]], "count", len(pnz))
    # random test vs torch
    B = 5
    dpre = torch.randn(B, T, Cout, device=dev); x = torch.randn(B, T, Cin, device=dev)
    G, db, _ = run(dpre, x, Cout, Cin, taps, pad)
    xp = torch.nn.functional.pad(x, (0, 0, pad, pad))
    ref = torch.stack([torch.einsum("bto,bti->oi", dpre, xp[:, j:j + T]) for j in range(taps)], 1).reshape(Cout, taps * Cin)
    print("random: rel err G", ((G - ref).norm() / ref.norm()).item(), "db", ((db - dpre.sum((0, 1))).norm() / dpre.sum((0, 1)).norm()).item())
