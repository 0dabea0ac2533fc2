"""How much of the critic's input-gradient error is explained by LeakyReLU kink flips alone?

The TemporalDiscriminator (seed-42 weights, eval-mode spectral norm) is run in float64 on CPU twice on the same
445 gestures: exactly, and with the backward mask of every LeakyReLU taken from a forward whose pre-activations carry a
relative error eps (what an fp32-grade / 3xTF32 / TF32 implementation has) - the arithmetic itself stays exact.  The
rel-L2 difference of d mean(D(x)) / dx between the two is the error floor ANY implementation with that forward accuracy
shows against fp64.   python scripts/kink_flip_sim.py > profiles/r02_kink_flip_floor.json"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.nn.functional as F
from oracle import torch_port
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from golden_util import Golden, MODS

class NoisyLeaky(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, eps, gen):
        noise = torch.randn(x.shape, generator=gen, dtype=x.dtype) * eps * x.abs().max()
        ctx.save_for_backward((x + noise) > 0)
        return F.leaky_relu(x, 0.2)
    @staticmethod
    def backward(ctx, g):
        (m,) = ctx.saved_tensors
        return g * torch.where(m, 1.0, 0.2), None, None

def run(D, x, eps, gen):
    x = x.clone().requires_grad_(True)
    h = x.transpose(1, 2)
    for m in D.temporal_conv:
        h = NoisyLeaky.apply(h, eps, gen) if isinstance(m, torch.nn.LeakyReLU) else m(h)
    h = F.adaptive_avg_pool1d(h, 8).flatten(1)
    for m in D.mlp:
        h = NoisyLeaky.apply(h, eps, gen) if isinstance(m, torch.nn.LeakyReLU) else m(h)
    D.output_layer(h).mean().backward()
    return x.grad

g = Golden("default")
tp = torch_port.TorchPortTrainer(seed=0, dtype=torch.float64)
tp.load_state({m: g.init_state(m) for m in MODS})
D = tp.D1.eval()
B = 445
x = torch.rand(B, 128, 3, generator=torch.Generator().manual_seed(5), dtype=torch.float64) * 2 - 1
ref = run(D, x, 0.0, torch.Generator().manual_seed(0))
out = {"what": __doc__.split("\n")[0], "B": B, "floor_rel_l2_of_dx_by_forward_rel_error": {}}
for eps in (1e-7, 1e-6, 2e-6, 1e-5, 1e-4, 1e-3):
    errs = []
    for s in range(3):
        d = run(D, x, eps, torch.Generator().manual_seed(10 + s))
        errs.append(float((d - ref).norm() / ref.norm()))
    out["floor_rel_l2_of_dx_by_forward_rel_error"][f"{eps:g}"] = errs
print(json.dumps(out, indent=1))
