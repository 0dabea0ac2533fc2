"""One warm-up + N timed training steps (for ncu launch lists and quick timing)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import wgg_b200 as wgg
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
mode = sys.argv[2] if len(sys.argv) > 2 else "tf32"
nsteps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
H = int(sys.argv[4]) if len(sys.argv) > 4 else 48      # gen_hidden_dim (scaled regime: 128 ... 1024)
T = int(sys.argv[5]) if len(sys.argv) > 5 else 128     # seq_length
dev = torch.device("cuda:0")
wgg.set_math_mode(mode)
wgg.seed_everything(42)
tr = wgg.WordGestureGANTrainer(wgg.ModelConfig(gen_hidden_dim=H, seq_length=T), wgg.TrainingConfig(), dev)
for m in (tr.generator, tr.encoder, tr.discriminator_1, tr.discriminator_2): m.train()
real = torch.rand(B, T, 3, device=dev) * 2 - 1
proto = torch.rand(B, T, 3, device=dev) * 2 - 1
wgg.train_batch(tr, real, proto, 1.0)
torch.cuda.synchronize()
t0 = time.perf_counter()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
torch.cuda.profiler.start()   # `ncu --profile-from-start off` captures exactly the timed steps
for _ in range(nsteps): wgg.train_batch(tr, real, proto, 1.0)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
t_issue = time.perf_counter() - t0
e1.record(); torch.cuda.synchronize()
print(f"B={B} mode={mode} H={H} T={T}: {e0.elapsed_time(e1)/nsteps:.2f} ms/step device, CPU issue time {t_issue/nsteps*1e3:.2f} ms/step")
