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
DEFAULT_MATH = "tf32"
PROFILE_CANDIDATES = ("gemm_kernel", "lstm_tc_fwd_kernel", "lstm_tc_bwd_kernel", "lstm_tc_dw_kernel", "lstm_tc_dx_kernel",
                      "conv_tc_fwd_kernel", "conv_tc_wgrad_kernel", "lstm_rec_fwd_kernel", "lstm_rec_bwd_kernel")


# scaled regime (--hidden > 64): input projections / weight and input gradients (tcgen05 GEMM), the persistent H = 128
# recurrence of the no-grad passes, the fused per-timestep recurrence (forward with stash / BPTT), the critics' conv kernels,
# operand transposes, the layer-0 projection, and what stays on the mma.sync / FMA engine
SCALED_CANDIDATES = ("gemm_kernel", "gemm_tc_nt_kernel", "lstm128_tc_fwd_kernel", "gemm_tc_lstm_fwd_kernel",
                     "gemm_tc_lstm_bwd_kernel", "conv_tc_fwd_kernel", "conv_tc_wgrad_kernel", "transpose_tf32_kernel",
                     "xproj0_kernel", "lstm_cell")


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


def synthetic_batch(batch, seed=0, T=128):
    import torch
    g = torch.Generator().manual_seed(seed)
    real = torch.rand(batch, T, 3, generator=g) * 2 - 1
    proto = torch.rand(batch, T, 3, generator=g) * 2 - 1
    return real, proto


def make_cpu_stepper(batch: int, threads: int):
    """The reference's CPU implementation of the step on all host threads: the reference's OWN files when they are
    reachable (/root/reference in the build container, the vendored oracle/_ref on the GPU box - oracle/ref_loader.py),
    else the torch.nn restatement oracle/torch_port.py.  Returns (kind, description, step_fn)."""
    import torch
    torch.set_num_threads(threads)
    real, proto = synthetic_batch(batch)
    try:
        from oracle.ref_runner import ReferenceRunner, reference_available
        if reference_available():
            rr = ReferenceRunner("cpu", seed=42, batch_size=batch)
            batches = [{"gesture": real, "prototype": proto}]
            return ("reference", "the unmodified reference: WordGestureGANTrainer + train_epoch_with_grad_clip "
                    f"(src/shared/utils.py:28) from {rr.ref.root}, torch CPU kernels (oneDNN LSTM/conv)",
                    lambda: rr.train_batches(batches))
    except Exception as ex:  # fall back to the restatement, say why
        print(f"[bench] reference not usable ({ex!r}); timing oracle/torch_port.py instead", file=sys.stderr)
    from oracle import torch_port
    tp = torch_port.TorchPortTrainer(seed=42)
    return ("port", "oracle/torch_port.py (the reference's modules restated on torch.nn; torch CPU kernels)",
            lambda: tp.train_batch(real, proto))


def cpu_baseline_gestures_per_s(batch: int, threads: int):
    kind, desc, step = make_cpu_stepper(batch, threads)
    step()  # warm-up
    t0 = time.perf_counter()
    step()
    dt = time.perf_counter() - t0
    return batch / dt, dt, kind, desc


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    batch = args.ref_batch
    kind, desc, step = make_cpu_stepper(batch, cores)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    value = batch / dt
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "impl": "reference",
        "config": {"workload": "default WordGesture-GAN train step (n_critic=5, TemporalDiscriminator x2, H=48 L=4 "
                               f"T=128), CPU, bounded sample of {batch} gestures per step", "batch_per_step": batch},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"{args.steps} steps x {batch} gestures, {cores} threads; {desc}"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def reference_cuda_leg(dev, batches, steps, sample_batch, model_kwargs=None):
    """The reference's library path on THIS GPU (SURVEY.md 8(d) last row): the unmodified reference trainer + epoch
    function on device cuda with torch's defaults, i.e. cuDNN LSTM / conv with TF32 allowed and fp32 cuBLAS
    (src/gan/models.py:160,163,270-291), its 11 .item() syncs per batch included (utils.py:84-131); the torch.nn
    restatement oracle/torch_port.py where the reference's files are not reachable.  Reported next to our numbers as
    the library-kernel bar; it is not the CPU baseline and not the driver's reference arm."""
    import torch
    out = {}
    try:
        from oracle.ref_runner import ReferenceRunner, reference_available
        use_ref = reference_available()
    except Exception:
        use_ref = False
    out["what"] = (("the unmodified reference (WordGestureGANTrainer + train_epoch_with_grad_clip)" if use_ref else
                    "oracle/torch_port.py") + " on cuda: cuDNN TF32 LSTM/conv + fp32 cuBLAS (torch defaults), eager, same step")
    try:
        for B in batches:
            T = (model_kwargs or {}).get("seq_length", 128)
            real, proto = synthetic_batch(B, T=T)
            real, proto = real.to(dev), proto.to(dev)
            if use_ref:
                rr = ReferenceRunner(dev, seed=42, batch_size=B, model_kwargs=model_kwargs)
                batch_list = [{"gesture": real, "prototype": proto}]
                step = lambda: rr.train_batches(batch_list)
                sample = rr.sample
            else:
                from oracle import torch_port
                from oracle.wgg_oracle import ModelCfg
                tp = torch_port.TorchPortTrainer(seed=42, device=dev, cfg=ModelCfg(**(model_kwargs or {})))
                step = lambda: tp.train_batch(real, proto)
                sample = tp.sample
            for _ in range(2):
                step()
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                step()
            e1.record()
            torch.cuda.synchronize(dev)
            ms = e0.elapsed_time(e1) / steps
            out[f"train_B{B}"] = {"value": B / (ms / 1e3), "unit": UNIT, "ms_per_step": ms, "steps": steps}
            if B == batches[-1]:
                g = torch.Generator().manual_seed(1)
                proto_s = (torch.rand(sample_batch, T, 3, generator=g) * 2 - 1).to(dev)
                z = torch.randn(sample_batch, 32, device=dev)
                for _ in range(2):
                    sample(proto_s, z)
                torch.cuda.synchronize(dev)
                e0.record()
                for _ in range(5):
                    sample(proto_s, z)
                e1.record()
                torch.cuda.synchronize(dev)
                out["sampling"] = {"value": sample_batch * 5 / (e0.elapsed_time(e1) / 1e3), "unit": "samples/s",
                                   "batch": sample_batch}
    except Exception as ex:  # a reported extra: never lose the GPU line because of it
        out["error"] = repr(ex)
    return out


def load_json(rel):
    path = os.path.join(ROOT, rel)
    if not os.path.exists(path):
        return None
    try:
        return json.load(open(path))
    except Exception:
        return None


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
    ap.add_argument("--no-reference-cuda", action="store_true")
    ap.add_argument("--sample-batch", type=int, default=148 * 128,
                    help="gestures per generator sampling call (148 tiles x 2 directions = two full waves of the persistent kernel)")
    ap.add_argument("--sample-total", type=int, default=1_000_000, help="BASELINE configs[4]: total samples over all GPUs")
    ap.add_argument("--profile-kernel", default="auto")
    ap.add_argument("--no-graph", action="store_true", help="issue every launch from Python instead of replaying the captured CUDA graph")
    ap.add_argument("--hidden", type=int, default=48, help="gen_hidden_dim (BASELINE configs[3]: scaled sweep 128 ... 1024)")
    ap.add_argument("--seq", type=int, default=128, help="seq_length (configs[3]: 256-point gestures)")
    ap.add_argument("--math", default=DEFAULT_MATH, choices=["fp32", "tf32", "tf32x3"],
                    help="tf32: LSTM/conv contractions on TF32 tensor cores (the reference CUDA path's numerics); tf32x3: conv "
                         "contractions error-compensated (every gradient within the north star's 1e-3); fp32: FMA only")
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
    T = args.seq
    mc, tc = wgg.ModelConfig(gen_hidden_dim=args.hidden, seq_length=T), wgg.TrainingConfig(batch_size=B)
    scaled = (args.hidden, T) != (48, 128)
    wgg.seed_everything(42)
    tr = wgg.WordGestureGANTrainer(mc, tc, dev)
    tr.use_cuda_graph = not args.no_graph
    dp = parallel.DataParallelGAN(tr) if world > 1 else None
    g = torch.Generator().manual_seed(1000 + rank)
    real_h = (torch.rand(B, T, 3, generator=g) * 2 - 1).pin_memory()
    proto_h = (torch.rand(B, T, 3, generator=g) * 2 - 1).pin_memory()
    real_d, proto_d = real_h.to(dev), proto_h.to(dev)
    host_batch = [{"gesture": real_h, "prototype": proto_h}]  # what a DataLoader(pin_memory=True) yields (data.py:526-533)

    def epoch_from_host(nbatches):
        """The reference's entry point (utils.py:28) over pinned host batches: H2D copies, the step, and the read-back
        of the epoch means inside."""
        return wgg.train_epoch_with_grad_clip(tr, host_batch * nbatches, 1.0, mc, tc, dev)

    # warm-up THROUGH the public entry point: the first full batch captures the step's CUDA graph, the rest replay it
    epoch_from_host(max(args.warmup, 3))
    gs = None
    if not args.no_graph:
        gs = tr._graphed_steps[(B, 1.0, args.math)]
    # per-gesture-step FLOPs of this configuration (BASELINE.md section 3 formula; 1.340 GFLOP for the default model)
    H = args.hidden
    f_g = T * (160 * H * H + 16 * 34 * H + 12 * H)
    f_d = (T / 128.0) * (245760 + 5242880 + 1572864) + 65536 + 16384 + 128
    f_e = (T * 3 * 192 + 192 * 96 + 96 * 48 + 48 * 32 + 2 * 32 * 32) * 2
    flop_per_gesture_step = 16 * f_g + 74 * f_d + 9 * f_e

    def step_eager():
        return wgg.train_batch(tr, real_d, proto_d, 1.0)

    def step_resident():
        return gs(real_d, proto_d) if gs is not None else step_eager()

    def step_e2e():
        return epoch_from_host(1)   # one batch per call: every step pays its H2D copies and the D2H read of its losses

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
        for cand in (PROFILE_CANDIDATES if not scaled else SCALED_CANDIDATES):
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
    tr.generator.eval()
    BS = args.sample_batch
    gs_ = torch.Generator().manual_seed(2000 + rank)
    if scaled:
        BS = min(BS, 4 * B)
        args.sample_total = min(args.sample_total, 16 * BS)
    proto_s = (torch.rand(BS, T, 3, generator=gs_) * 2 - 1).to(dev)
    zs = torch.randn(BS, 32, device=dev)

    def sample():
        with torch.no_grad():
            return tr.generator(proto_s, zs)

    for _ in range(2):
        sample()
    ms_s = timed(sample, 10)
    samples_per_s = world * BS * 10 / (ms_s / 1e3)
    # BASELINE configs[4]: sample_total samples over all GPUs (each rank its shard, no collective), in calls of BS
    per_rank = args.sample_total // world
    calls = max(1, (per_rank + BS - 1) // BS)
    out_buf = torch.empty(min(per_rank, calls * BS), T, 3, device=dev)

    def sample_shard():
        with torch.no_grad():
            for c in range(calls):
                n = min(BS, per_rank - c * BS)
                if n <= 0:
                    break
                out_buf[c * BS:c * BS + n].copy_(tr.generator(proto_s[:n], zs[:n]))

    sample_shard()
    ms_1m = timed(sample_shard, 1)
    del out_buf

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
    # DRAM traffic of the same kernel class over the SAME launch mix (every launch of one step), from the committed
    # `ncu --set full` capture of this command (scripts/ncu_traffic.py writes the file); null if there is none
    traffic, traffic_src = None, None
    tr_file = load_json(f"profiles/r02_ncu_traffic_{prof_kernel}.json")
    if tr_file and tr_file.get("batch_per_gpu") == B and tr_file.get("math") == args.math:
        traffic = tr_file["dram_bytes_per_launch_mean"]
        traffic_src = (f"profiles/r02_ncu_traffic_{prof_kernel}.json: mean dram__bytes_read.sum + dram__bytes_write.sum over "
                       f"the {tr_file['launches']} launches of one step (same launch mix as algorithmic_bytes_per_launch)")
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
            "whole_step": {"achieved": flop_per_gesture_step * value / world / 1e12, "unit": "TFLOP/s",
                           "frac": flop_per_gesture_step * value / world / 1e12 / peaks["tf32_tflops"],
                           "flop_per_gesture_step": flop_per_gesture_step}}
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        # rank 0 at N = 1 only: under torchrun the other ranks would sit in an NCCL barrier (GPUs busy-waiting) while
        # this leg runs, and the host cores are shared by N processes
        cores = os.cpu_count() or 1
        try:
            v, dt, kind, desc = cpu_baseline_gestures_per_s(args.cpu_sample, cores)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
                   "sample": f"1 step x {args.cpu_sample} gestures ({dt:.1f} s) after 1 warm-up, {cores} threads; {desc}"}
        except Exception as ex:  # the baseline is a reported extra; never lose the GPU line because of it
            cpu = {"value": None, "unit": UNIT, "cores": cores, "kind": "port", "sample": f"failed: {ex!r}"}
    ref_cuda = None
    if world == 1 and not args.no_reference_cuda:
        ref_cuda = reference_cuda_leg(dev, (512, B) if B != 512 else (512,), 3, BS,
                                      dict(gen_hidden_dim=args.hidden, seq_length=T) if scaled else None)
    numerics = load_json("profiles/r02_numerics_by_mode.json")
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"tf32": "tf32", "tf32x3": "tf32x3", "fp32": "f32"}[args.math],
        "math_mode": args.math,
        "numerics": (numerics or {}).get(args.math, {"note": "profiles/r02_numerics_by_mode.json not found"}),
        "data": "synthetic",
        "config": {"workload": ("default WordGesture-GAN train step (n_critic=5, TemporalDiscriminator x2, H=48 L=4 "
                                "T=128), BASELINE configs[1]") if not scaled else
                               (f"scaled WordGesture-GAN train step (BASELINE configs[3]): gen_hidden_dim={H}, seq_length={T}, "
                                "n_critic=5, TemporalDiscriminator x2, L=4"),
                   "batch_per_gpu": B, "global_batch": B * world,
                   "parallelism": f"dp{world}", "launch": "eager" if gs is None else "cuda-graph (1 replay per step)",
                   "math_mode": args.math,
                   "gradient_exchange": None if dp is None else ("one-shot peer-memory reduce (wgg_p2p_allreduce_avg), 12 per step"
                                                                  if dp.p2p else "NCCL all-reduce (AVG), 11 per step"),
                   "l2": "per-step working set (GBs of activations) >> 126 MB L2; no flush needed"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 2 * B * 128 * 3 * 4,
                "d2h_bytes_per_step": 16 + 4, "ms_per_step": ms_e2e / args.steps,
                "api": "train_epoch_with_grad_clip(trainer, [pinned CPU batch dict], 1.0, model_config, training_config, device) "
                       "once per step: H2D of gesture + prototype, one graph replay, D2H of the four loss means + the "
                       "pipeline-health word"},
        "gpu_launches": int(launches),
        "roofline": roof,
        "cpu_baseline": cpu,
        "reference_cuda": ref_cuda,
        "clocks": clocks,
        "sampling": {"value": samples_per_s, "unit": "samples/s", "batch_per_gpu": BS,
                     "roofline_frac": f_g * samples_per_s / world / 1e12 / peaks["tf32_tflops"],
                     "configs4": {"total_samples": per_rank * world, "per_gpu": per_rank, "ms": ms_1m,
                                  "value": per_rank * world / (ms_1m / 1e3), "unit": "samples/s",
                                  "note": "BASELINE configs[4]: every rank generates its shard in calls of batch_per_gpu "
                                          "and writes it to one HBM buffer; no collective; max over ranks"}},
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
