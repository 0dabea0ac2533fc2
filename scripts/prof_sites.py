"""Per-call-site device-time breakdown of one training step (CUDA-event brackets inside libwgg_sm100)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import wgg_b200 as wgg
from wgg_b200 import _lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
mode = sys.argv[2] if len(sys.argv) > 2 else "tf32"
H = int(sys.argv[3]) if len(sys.argv) > 3 else 48
T = int(sys.argv[4]) if len(sys.argv) > 4 else 128
dev = torch.device("cuda:0")
wgg.set_math_mode(mode)
wgg.seed_everything(42)
tr = wgg.WordGestureGANTrainer(wgg.ModelConfig(gen_hidden_dim=H, seq_length=T), wgg.TrainingConfig(), dev)
for m in (tr.generator, tr.encoder, tr.discriminator_1, tr.discriminator_2): m.train()
real = torch.rand(B, T, 3, device=dev) * 2 - 1
proto = torch.rand(B, T, 3, device=dev) * 2 - 1
for _ in range(2): wgg.train_batch(tr, real, proto, 1.0)
torch.cuda.synchronize()
out = {}
for filt in ("gemm_kernel", "lstm_", "kernel"):
    _lib.profile_enable(dev, filt)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); wgg.train_batch(tr, real, proto, 1.0); e1.record(); torch.cuda.synchronize()
    rows = _lib.profile_report(dev)
    out[filt] = dict(step_ms=e0.elapsed_time(e1), rows=rows)
    print(f"--- filter {filt}: step {e0.elapsed_time(e1):.1f} ms")
    for r in sorted(rows, key=lambda r: -r["ms"]):
        tf = r["gflop"] / r["ms"] if r["ms"] > 0 else 0
        print(f"{r['tag']:34s} n={r['launches']:5d} {r['ms']:9.2f} ms {r['gflop']:10.1f} GFLOP {tf:8.2f} TFLOP/s")
    _lib.profile_enable(dev, None)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open(f"gpurun_out/prof_sites_{mode}_{B}_H{H}_T{T}.json", "w"), indent=1)
