"""How far is the reference's own CUDA library path (cuDNN TF32 conv/LSTM, the torch defaults) from fp64?
Uses oracle/torch_port.py modules on cuda vs the same modules in fp64 on CPU.  Reported per tensor."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import torch_port
from oracle.wgg_oracle import ModelCfg

def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-300)).item()

torch.manual_seed(0)
cfg = ModelCfg()
B = 24
ref = torch_port.TorchPortTrainer(seed=42, cfg=cfg, dtype=torch.float64)
states = ref.state()
report = {}
for tf32 in (True, False):
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    gpu = torch_port.TorchPortTrainer(seed=42, cfg=cfg, dtype=torch.float32)
    gpu.load_state(states)
    for m in gpu.mods.values(): m.cuda()
    ref.load_state(states)
    g = torch.Generator().manual_seed(1)
    real = (torch.rand(B, 128, 3, generator=g) * 2 - 1)
    proto = (torch.rand(B, 128, 3, generator=g) * 2 - 1)
    z = torch.randn(B, 32, generator=g)
    r = {}
    # discriminator: -mean(D(x)) gradients wrt params and input
    for name in ("D1",):
        xr = real.double().requires_grad_(True); xg = real.cuda().requires_grad_(True)
        for mod in (ref.mods[name], gpu.mods[name]): mod.zero_grad()
        (-ref.mods[name](xr).mean()).backward(); (-gpu.mods[name](xg).mean()).backward()
        r["disc_dx"] = rel(xg.grad, xr.grad)
        for (k, pr), (_, pg) in zip(ref.mods[name].named_parameters(), gpu.mods[name].named_parameters()):
            r["disc_grad/" + k] = rel(pg.grad, pr.grad)
    # generator
    zr = z.double().requires_grad_(True); zg = z.cuda().requires_grad_(True)
    dy = torch.randn(B, 128, 3, generator=g)
    ref.G.zero_grad(); gpu.G.zero_grad()
    yr = ref.G(proto.double(), zr); yg = gpu.G(proto.cuda(), zg)
    r["gen_fwd_maxabs"] = ((yg.double().cpu() - yr).abs().max() / yr.abs().max()).item()
    yr.backward(dy.double()); yg.backward(dy.cuda())
    r["gen_dz"] = rel(zg.grad, zr.grad)
    r["gen_grad_worst"] = max(rel(pg.grad, pr.grad) for (_, pr), (_, pg) in zip(ref.G.named_parameters(), gpu.G.named_parameters()))
    report["cudnn_tf32" if tf32 else "cudnn_fp32"] = r
os.makedirs("gpurun_out", exist_ok=True)
json.dump(report, open("gpurun_out/ref_cuda_precision.json", "w"), indent=1)
for k, r in report.items():
    print(k)
    for n, v in r.items(): print(f"   {n:45s} {v:.3e}")
