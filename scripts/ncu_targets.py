"""One pass over every heavy kernel class (for `ncu --profile-from-start off`): a grad-carrying generator forward +
backward at 8192 gestures, one critic step of D1 (real + fake, 4096 gestures each) and one feature-matching loss
forward + backward.  Warm-up first; the profiled region is bracketed with cudaProfilerStart/Stop."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import wgg_b200 as wgg
from wgg_b200.gan_losses import WassersteinLoss
dev = torch.device("cuda:0")
wgg.set_math_mode("tf32"); wgg.seed_everything(42)
tr = wgg.WordGestureGANTrainer(wgg.ModelConfig(), wgg.TrainingConfig(), dev)
for m in (tr.generator, tr.encoder, tr.discriminator_1, tr.discriminator_2): m.train()
B = 4096
real = torch.rand(B, 128, 3, device=dev) * 2 - 1
proto2 = torch.rand(2 * B, 128, 3, device=dev) * 2 - 1
z2 = torch.randn(2 * B, 32, device=dev)

def region():
    # generator forward (stash) + backward
    tr.optimizer_G.zero_grad()
    fake = tr.generator(proto2, z2)
    fake.square().mean().backward()
    # one critic step of D1
    tr.optimizer_D1.zero_grad()
    f = fake.detach()[:B]
    loss = WassersteinLoss.discriminator_loss(tr.discriminator_1(real), tr.discriminator_1(f))
    loss.backward()
    tr.optimizer_D1.step(max_norm=1.0)
    # feature matching forward + backward through D2 (input gradient path: dgrad / dx kernels)
    f2 = fake.detach()[B:].requires_grad_(True)
    wgan, feat = tr._adversarial_terms(tr.discriminator_2, f2, real)
    (wgan + feat).backward()

region(); torch.cuda.synchronize()
torch.cuda.profiler.start()
region(); torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
