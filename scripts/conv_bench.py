"""Micro-benchmark of the tcgen05 conv weight-gradient kernel per layer shape (B200)."""
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
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
st = torch.cuda.current_stream().cuda_stream
for (Cout, Cin, taps, pad) in [(64, 64, 5, 2), (32, 64, 3, 1), (64, 3, 5, 2)]:
    Cin4 = (Cin + 3) // 4 * 4
    dpc = torch.randn(B, Cout // 4, T, 4, device=dev); xc = torch.randn(B, Cin4 // 4, T, 4, device=dev)
    G = torch.zeros(Cout, taps * Cin, device=dev); db = torch.zeros(Cout, device=dev)
    ws = torch.zeros(256 * 64 * 336, device=dev)
    def call():
        rc = lib.wgg_debug_conv_tc_wgrad(c, dpc.data_ptr(), xc.data_ptr(), B, Cout, Cin, taps, pad, G.data_ptr(), db.data_ptr(), ws.data_ptr(), st)
        assert rc == 0
    for _ in range(5): call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): call()
    e1.record(); torch.cuda.synchronize()
    tot = e0.elapsed_time(e1) / 20 * 1e3
    _lib.profile_enable(dev, "conv_tc_wgrad_kernel")
    for _ in range(20): call()
    pr = _lib.profile_read(dev); _lib.profile_enable(dev, None)
    k_us = pr["ms"] / max(pr["launches"], 1) * 1e3
    nbytes = B * T * 4 * (Cout + Cin4)
    print(f"wgrad Cout={Cout} Cin={Cin} taps={taps}: wgrad+finalize {tot:7.1f} us; kernel {k_us:7.1f} us -> {nbytes / k_us / 1e3:7.1f} GB/s "
          f"(floor {nbytes / 6.55e3 / 1e3:5.1f} us), async_err {_lib.async_error(dev)}")

