"""No-grad generator forward in the scaled regime (for ncu captures of gemm_tc_nt_kernel and quick timing).
usage: scaled_forward.py H T B"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import wgg_b200 as wgg
from wgg_b200 import _lib
H, T, B = (int(a) for a in sys.argv[1:4])
dev = torch.device("cuda:0")
wgg.set_math_mode("tf32"); wgg.seed_everything(42)
G = wgg.Generator(wgg.ModelConfig(gen_hidden_dim=H, seq_length=T)).to(dev).eval()
proto = torch.rand(B, T, 3, device=dev) * 2 - 1
z = torch.randn(B, 32, device=dev)
with torch.no_grad():
    y = G(proto, z)
    torch.cuda.synchronize()
    _lib.profile_enable(dev, "gemm_tc_nt_kernel")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); y = G(proto, z); e1.record(); torch.cuda.synchronize()
    pr = _lib.profile_read(dev)
flops_fwd = B * T * (160 * H * H + 16 * 34 * H + 12 * H)
print(f"H={H} T={T} B={B}: forward {e0.elapsed_time(e1):.2f} ms = {flops_fwd / e0.elapsed_time(e1) / 1e9:.1f} TFLOP/s whole forward; "
      f"gemm_tc_nt_kernel: {pr['launches']} launches, {pr['ms']:.2f} ms, {pr['flops'] / max(pr['ms'], 1e-9) / 1e9:.1f} TFLOP/s; async_err {_lib.async_error(dev)}")
