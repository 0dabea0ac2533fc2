"""Per-tensor error of our discriminator path vs the fp64 oracle in each math mode."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import wgg_b200 as wgg
from oracle import wgg_oracle as O
from gpu_util import DEV, grads_of, model_cfg, rand_inputs, state_of, to_np, to_t
from golden_util import rel_l2
ocfg = O.ModelCfg(); B = 24
for mode in ("fp32", "tf32"):
    wgg.set_math_mode(mode)
    torch.manual_seed(3)
    D = wgg.TemporalDiscriminator(model_cfg(ocfg)).to(DEV).train()
    pd = state_of(D)
    real, _, _ = rand_inputs(ocfg, B, 3)
    rs_ref, _, st_r = O.disc_fwd(pd, ocfg, real, True)
    g_r, dx_ref = O.disc_bwd(pd, ocfg, st_r, np.full((B, 1), -1.0 / B), None)
    xt = to_t(real).requires_grad_(True)
    wgg.WassersteinLoss.generator_loss(D(xt)).backward()
    print(mode, "disc_dx", f"{rel_l2(to_np(xt.grad), dx_ref):.3e}")
    for k, v in grads_of(D).items(): print(f"   {mode} disc_grad/{k:40s} {rel_l2(v, g_r[k]):.3e}")
