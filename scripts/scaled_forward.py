"""No-grad generator forward in the scaled regime (BASELINE configs[3]) - quick timing and the target of ncu captures.
Reports the eager and the CUDA-graph forward time and, per tcgen05 kernel class, launches / time / TFLOP/s
(gemm_tc_nt_kernel: the input projections; gemm_tc_lstm_fwd_kernel: the fused per-timestep recurrence).
usage: scaled_forward.py H T B [nograph]"""
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
flops_fwd = B * T * (160 * H * H + 16 * 34 * H + 12 * H)
ev = lambda: torch.cuda.Event(enable_timing=True)
with torch.no_grad():
    y = G(proto, z)
    torch.cuda.synchronize()
    e0, e1 = ev(), ev()
    e0.record(); y = G(proto, z); e1.record(); torch.cuda.synchronize()
    eager = e0.elapsed_time(e1)
    parts = []
    for k in ("gemm_tc_nt_kernel", "gemm_tc_lstm_fwd_kernel", "lstm128_tc_fwd_kernel", "xproj0_kernel"):
        _lib.profile_enable(dev, k)
        G(proto, z); torch.cuda.synchronize()
        pr = _lib.profile_read(dev)
        parts.append(f"{k}: {pr['launches']} launches, {pr['ms']:.2f} ms, {pr['flops'] / max(pr['ms'], 1e-9) / 1e9:.1f} TFLOP/s")
    _lib.profile_enable(dev, None)
    graphed = float("nan")
    if len(sys.argv) < 5:
        scope = _lib.ScratchScope()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), scope:
            G(proto, z)
        torch.cuda.current_stream(dev).wait_stream(side); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        scope.frozen = True
        with scope, torch.cuda.graph(g):
            y = G(proto, z)
        g.replay(); torch.cuda.synchronize()
        e0, e1 = ev(), ev()
        e0.record(); g.replay(); g.replay(); g.replay(); e1.record(); torch.cuda.synchronize()
        graphed = e0.elapsed_time(e1) / 3
print(f"H={H} T={T} B={B}: forward eager {eager:.2f} ms, graph {graphed:.2f} ms = {flops_fwd / min(eager, graphed if graphed == graphed else eager) / 1e9:.1f} TFLOP/s whole forward; "
      + "; ".join(parts) + f"; async_err {_lib.async_error(dev)}")
