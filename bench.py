#!/usr/bin/env python
"""Benchmark of the WordGesture-GAN training step (BASELINE.json metric: GAN train gestures/sec, G+D step).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]

A "step" is one DataLoader batch of train_epoch_with_grad_clip: n_critic x (D1 step + D2 step) + one joint
G/E step (src/shared/utils.py:62-135) on the default model.  Workload = BASELINE.json configs[1]: default model,
4096 synthetic gestures of dataset shape (128 x 3) per GPU, fp32.  Under torchrun each rank owns its own
4096-gesture shard (weak scaling) and gradients are mean-all-reduced over NCCL before every optimiser step.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "GAN train gestures/sec (G+D step)"
UNIT = "gestures/s"
FLOP_PER_GESTURE_STEP = 1.340e9  # BASELINE.md section 3 (default model, as-written call counts)


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm_gbs=p["hbm_gbs"], tf32_tflops=p["bf16_tflops_sustained"] / 2.0, source="measured (MEASURED_PEAKS.json; "
                    "TF32 dense = 1/2 of the sustained bf16 cuBLAS figure)")
    return dict(hbm_gbs=6650.0, tf32_tflops=1400.0 / 2.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        mhz, mx, reasons = [], None, set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for s in self.samples:
            parts = [x.strip() for x in s.split(",")]
            if len(parts) < 6:
                continue
            try:
                mhz.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        mhz.sort()
        return {"sm_mhz": mhz[len(mhz) // 2] if mhz else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(mhz)}


def cpu_port_gestures_per_s(batch: int, threads: int):
    """Times the CPU restatement of the step (oracle/torch_port.py: the reference's modules restated with the same
    torch CPU library kernels the reference itself runs on, all host threads) on a bounded sample."""
    import torch
    from oracle import torch_port
    torch.set_num_threads(threads)
    tp = torch_port.TorchPortTrainer(seed=42)
    g = torch.Generator().manual_seed(0)
    real = torch.rand(batch, 128, 3, generator=g) * 2 - 1
    proto = torch.rand(batch, 128, 3, generator=g) * 2 - 1
    tp.train_batch(real, proto)  # warm-up
    t0 = time.perf_counter()
    tp.train_batch(real, proto)
    dt = time.perf_counter() - t0
    return batch / dt, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    cores = os.cpu_count() or 1
    vals = []
    t_all = time.perf_counter()
    batch = args.ref_batch
    from oracle import torch_port
    torch.set_num_threads(cores)
    tp = torch_port.TorchPortTrainer(seed=42)
    g = torch.Generator().manual_seed(0)
    real = torch.rand(batch, 128, 3, generator=g) * 2 - 1
    proto = torch.rand(batch, 128, 3, generator=g) * 2 - 1
    for _ in range(args.warmup):
        tp.train_batch(real, proto)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        tp.train_batch(real, proto)
    dt = (time.perf_counter() - t0) / args.steps
    value = batch / dt
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "impl": "reference",
        "config": {"workload": "default WordGesture-GAN train step (n_critic=5, TemporalDiscriminator x2, H=48 L=4 "
                               f"T=128), CPU, bounded sample of {batch} gestures per step", "batch_per_step": batch},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{args.steps} steps x {batch} gestures, torch CPU kernels (oneDNN LSTM/conv), "
                                   f"{cores} threads; the reference is pure Python and cannot travel to this box, so "
                                   "oracle/torch_port.py restates its modules on the same library path"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=4096, help="gestures per GPU per step")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-batch", type=int, default=512)
    ap.add_argument("--cpu-sample", type=int, default=512)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sample-batch", type=int, default=74 * 128, help="gestures per generator sampling call")
    ap.add_argument("--profile-kernel", default="auto")
    ap.add_argument("--no-graph", action="store_true", help="issue every launch from Python instead of replaying the captured CUDA graph")
    ap.add_argument("--math", default="tf32", choices=["fp32", "tf32", "tf32x3"],
                    help="tf32: LSTM/conv contractions on TF32 tensor cores (the reference CUDA path's numerics); fp32: FMA only")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    import wgg_b200 as wgg
    from wgg_b200 import _lib, parallel

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch
    wgg.set_math_mode(args.math)
    mc, tc = wgg.ModelConfig(), wgg.TrainingConfig(batch_size=B)
    wgg.seed_everything(42)
    tr = wgg.WordGestureGANTrainer(mc, tc, dev)
    if world > 1:
        parallel.DataParallelGAN(tr)
    for m in (tr.generator, tr.encoder, tr.discriminator_1, tr.discriminator_2):
        m.train()
    g = torch.Generator().manual_seed(1000 + rank)
    real_h = (torch.rand(B, 128, 3, generator=g) * 2 - 1).pin_memory()
    proto_h = (torch.rand(B, 128, 3, generator=g) * 2 - 1).pin_memory()
    real_d, proto_d = real_h.to(dev), proto_h.to(dev)
    keys = ("d1_loss", "d2_loss", "cycle1_total", "cycle2_total")

    gs = None if args.no_graph else wgg.GraphedTrainStep(tr, B, 1.0)

    def step_eager():
        return wgg.train_batch(tr, real_d, proto_d, 1.0)

    def step_resident():
        return gs(real_d, proto_d) if gs is not None else step_eager()

    def step_e2e():
        if gs is not None:
            out = gs(real_h, proto_h)  # pinned host -> the graph's static device inputs, then one replay
        else:
            out = wgg.train_batch(tr, real_h.to(dev, non_blocking=True), proto_h.to(dev, non_blocking=True), 1.0)
        return torch.stack([out[k] for k in keys]).tolist()  # D2H read of the step's losses

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps):
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        sync()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms

    for _ in range(args.warmup):
        step_resident()
    sync()
    # choose the kernel class to time: the one with the largest share of one step
    prof_kernel = args.profile_kernel
    shares = {}
    if prof_kernel == "auto":
        for cand in ("gemm_kernel", "lstm_tc_fwd_kernel", "lstm_tc_bwd_kernel", "lstm_tc_dw_kernel", "lstm_tc_dx_kernel",
                     "conv_tc_fwd_kernel", "conv_tc_wgrad_kernel", "lstm_rec_fwd_kernel", "lstm_rec_bwd_kernel"):
            _lib.profile_enable(dev, cand)
            step_eager()
            shares[cand] = _lib.profile_read(dev)["ms"]
        prof_kernel = max(shares, key=shares.get)
    _lib.profile_enable(dev, None)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = _lib.launch_count(dev)
    ms = timed(step_resident, args.steps)
    launches = (_lib.launch_count(dev) - l0) if gs is None else gs.launches_per_step * args.steps
    # second timed pass with event brackets around the dominant kernel class (kept separate so that the
    # brackets cannot perturb the headline number)
    _lib.profile_enable(dev, prof_kernel)
    prof_steps = min(args.steps, 4)
    ms_prof = timed(step_eager, prof_steps)  # brackets need eager launches (events cannot be recorded inside a replay)
    prof = _lib.profile_read(dev)
    _lib.profile_enable(dev, None)
    ms_e2e = timed(step_e2e, args.steps)
    clocks = sampler.stop() if rank == 0 else None

    # generator sampling throughput (second half of the BASELINE metric): eval / no-grad, output written to HBM.
    # Batch = 74 tiles x 128 gestures: with two directions that is exactly one wave of the persistent recurrent
    # kernel on 148 SMs.
    tr.generator.eval()
    BS = args.sample_batch
    gs_ = torch.Generator().manual_seed(2000 + rank)
    proto_s = (torch.rand(BS, 128, 3, generator=gs_) * 2 - 1).to(dev)
    zs = torch.randn(BS, 32, device=dev)

    def sample():
        with torch.no_grad():
            return tr.generator(proto_s, zs)

    for _ in range(2):
        sample()
    ms_s = timed(sample, 10)
    samples_per_s = world * BS * 10 / (ms_s / 1e3)

    if rank != 0:
        return finish(world, dev)
    peaks = load_peaks()
    ms_step = ms / args.steps
    value = world * B / (ms_step / 1e3)
    e2e_value = world * B / (ms_e2e / args.steps / 1e3)
    k_ms = prof["ms"] / max(prof["launches"], 1)
    k_tflops = prof["flops"] / max(prof["ms"], 1e-9) / 1e9
    k_gbs = prof["bytes"] / max(prof["ms"], 1e-9) / 1e6
    # which roof bounds the dominant kernel: its arithmetic intensity against the ridge of the two measured peaks
    ai = prof["flops"] / max(prof["bytes"], 1.0)
    ridge = peaks["tf32_tflops"] * 1e12 / (peaks["hbm_gbs"] * 1e9)
    hbm_bound = ai < ridge
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "r01_ncu_full_lstm_tc_fwd_B4096.json")
    if prof_kernel == "lstm_tc_fwd_kernel" and os.path.exists(tpath):
        try:
            caps = json.load(open(tpath))
            def gb(c, key):
                return float(next(v for k, v in c.items() if k.startswith(key)).replace(",", ""))
            traffic = sum(gb(c, "dram__bytes_read.sum") + gb(c, "dram__bytes_write.sum") for c in caps) / len(caps) * 1e9
            traffic_src = ("profiles/r01_ncu_full_lstm_tc_fwd_B4096.json: mean DRAM read+write bytes of the captured launches "
                           "(the three 40960-gesture critic-phase layer launches, B=4096; algorithmic 3.6 / 6.0 / 6.0 GB)")
        except Exception:
            traffic = None
    roof = {"bound": "hbm" if hbm_bound else "tensor", "kernel": prof_kernel,
            "achieved": k_gbs if hbm_bound else k_tflops, "peak": peaks["hbm_gbs"] if hbm_bound else peaks["tf32_tflops"],
            "unit": "GB/s" if hbm_bound else "TFLOP/s",
            "frac": (k_gbs / peaks["hbm_gbs"]) if hbm_bound else (k_tflops / peaks["tf32_tflops"]),
            "traffic": traffic, "traffic_source": traffic_src,
            "arithmetic_intensity_flop_per_byte": ai, "ridge_flop_per_byte": ridge,
            "tensor": {"achieved": k_tflops, "peak": peaks["tf32_tflops"], "unit": "TFLOP/s", "frac": k_tflops / peaks["tf32_tflops"]},
            "hbm": {"achieved": k_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": k_gbs / peaks["hbm_gbs"]},
            "launches_per_step": prof["launches"] / prof_steps, "avg_launch_ms": k_ms,
            "algorithmic_bytes_per_launch": prof["bytes"] / max(prof["launches"], 1),
            "share_of_step": prof["ms"] / max(ms_prof, 1e-9), "peak_source": peaks["source"],
            "kernel_share_ms_per_step": shares,
            "whole_step": {"achieved": FLOP_PER_GESTURE_STEP * value / world / 1e12, "unit": "TFLOP/s",
                           "frac": FLOP_PER_GESTURE_STEP * value / world / 1e12 / peaks["tf32_tflops"]}}
    cpu = None
    if not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        try:
            v, dt = cpu_port_gestures_per_s(args.cpu_sample, cores)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"1 step x {args.cpu_sample} gestures ({dt:.1f} s) after 1 warm-up, oracle/torch_port.py "
                             f"(torch CPU kernels, {cores} threads)"}
        except Exception as ex:  # the baseline is a reported extra; never lose the GPU line because of it
            cpu = {"value": None, "unit": UNIT, "cores": cores, "kind": "port", "sample": f"failed: {ex!r}"}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"tf32": "tf32", "tf32x3": "tf32x3", "fp32": "f32"}[args.math],
        "data": "synthetic",
        "config": {"workload": "default WordGesture-GAN train step (n_critic=5, TemporalDiscriminator x2, H=48 L=4 "
                               "T=128), BASELINE configs[1]", "batch_per_gpu": B, "global_batch": B * world,
                   "parallelism": f"dp{world}", "launch": "eager" if gs is None else "cuda-graph (1 replay per step)", "l2": "per-step working set (GBs of activations) >> 126 MB L2; no flush needed"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 2 * B * 128 * 3 * 4,
                "d2h_bytes_per_step": 16, "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
        "roofline": roof,
        "cpu_baseline": cpu,
        "clocks": clocks,
        "sampling": {"value": samples_per_s, "unit": "samples/s", "batch_per_gpu": BS,
                     "roofline_frac": 50.6e6 * samples_per_s / world / 1e12 / peaks["tf32_tflops"]},
    }
    print(json.dumps(line), flush=True)
    finish(world, dev)


def finish(world, dev):
    """Multi-rank exit: tearing an NCCL communicator down while captured graphs still reference it can block, so
    ranks meet at a barrier and leave without the teardown (the process is ending anyway)."""
    if world <= 1:
        return
    import torch
    import torch.distributed as dist
    torch.cuda.synchronize(dev)
    dist.barrier()
    torch.cuda.synchronize(dev)
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
