"""-m gpu: the scaled-model regime (BASELINE configs[3]: gen_hidden_dim 128 ... 1024, 256-point gestures).
``gen_hidden_dim`` / ``seq_length`` are free knobs of the reference (src/shared/config.py:15,22; src/gan/models.py:114-120);
beyond H = 64 the recurrence runs step by step on the contraction engine (csrc/lstm.cu: rec_fwd_generic / rec_bwd_generic).
Checked against the fp64 oracles with the tolerances of tests/test_gpu_parity.py (fp32: forward 1e-4, gradients 1e-3) and
of the TF32 modes (forward 2e-3, generator gradients 2e-3)."""
import numpy as np
import pytest
import torch

import wgg_b200 as wgg
from golden_util import LOSS_KEYS, max_abs_rel, rel_l2
from gpu_util import ATTR, DEV, grads_of, load_state, model_cfg, rand_inputs, state_of, to_np, to_t
from oracle import torch_port
from oracle import wgg_oracle as O

pytestmark = pytest.mark.gpu

H128 = O.ModelCfg(seq_length=256, gen_hidden_dim=128, gen_num_layers=2)
H96 = O.ModelCfg(seq_length=64, gen_hidden_dim=96, gen_num_layers=3)
H256 = O.ModelCfg(seq_length=32, gen_hidden_dim=256, gen_num_layers=1)


@pytest.fixture(params=["fp32", "tf32"])
def mode(request):
    wgg.set_math_mode(request.param)
    yield request.param
    wgg.set_math_mode("fp32")


@pytest.mark.parametrize("ocfg,B,seed", [(H128, 5, 0), (H96, 3, 1), (H256, 9, 2)])
def test_generator_any_hidden_size(mode, ocfg, B, seed):
    torch.manual_seed(seed)
    G = wgg.Generator(model_cfg(ocfg)).to(DEV)
    p = state_of(G)
    _, proto, z = rand_inputs(ocfg, B, seed)
    dy = np.random.default_rng(seed).standard_normal((B, ocfg.seq_length, 3)).astype(np.float32).astype(np.float64)
    y_ref, stash = O.generator_fwd(p, ocfg, proto, z)
    g_ref, dz_ref = O.generator_bwd(p, ocfg, stash, dy)
    zt = to_t(z).requires_grad_(True)
    y = G(to_t(proto), zt)
    y.backward(to_t(dy))
    with torch.no_grad():
        y_ng = G(to_t(proto), to_t(z))
    fwd_tol, grad_tol = (1e-4, 1e-3) if mode == "fp32" else (2e-3, 2e-3)
    assert max_abs_rel(to_np(y), y_ref) <= fwd_tol and max_abs_rel(to_np(y_ng), y_ref) <= fwd_tol
    worst = max(rel_l2(v, g_ref[k]) for k, v in grads_of(G).items())
    e_dz = rel_l2(to_np(zt.grad), dz_ref)
    print(f"{mode} H={ocfg.gen_hidden_dim} T={ocfg.seq_length}: fwd {max_abs_rel(to_np(y), y_ref):.2e} grads {worst:.2e} dz {e_dz:.2e}")
    assert worst <= grad_tol and e_dz <= grad_tol, (worst, e_dz)


@pytest.mark.parametrize("H,B,T,L", [(128, 300, 12, 2), (256, 129, 6, 2), (64, 520, 10, 2), (128, 260, 9, 3), (256, 132, 8, 2)])
def test_generator_large_batch_tcgen05_gemm(H, B, T, L):
    """Batches large enough (M >= 128) for the TMA-fed tcgen05 GEMM (csrc/gemm_tc.cu) to take the input projection and
    the per-step recurrent product in mode tf32: ragged M tiles (B not a multiple of 256), N = 4H in {256, 512, 1024},
    K in {64 ... 512}; forward and gradients against the fp64 oracle.  With B a multiple of 4 the weight and input
    gradients of layers >= 1 also run on it (split-K over K = T * B, transposed TF32 operand images, csrc/lstm.cu)."""
    from wgg_b200 import _lib
    ocfg = O.ModelCfg(seq_length=T, gen_hidden_dim=H, gen_num_layers=L)
    wgg.set_math_mode("tf32")
    try:
        torch.manual_seed(H)
        G = wgg.Generator(model_cfg(ocfg)).to(DEV)
        p = state_of(G)
        _, proto, z = rand_inputs(ocfg, B, H)
        dy = np.random.default_rng(H).standard_normal((B, T, 3)).astype(np.float32).astype(np.float64)
        y_ref, stash = O.generator_fwd(p, ocfg, proto, z)
        g_ref, dz_ref = O.generator_bwd(p, ocfg, stash, dy)
        _lib.profile_enable(DEV, "tc_")
        zt = to_t(z).requires_grad_(True)
        y = G(to_t(proto), zt)
        fwd_rows = {r["tag"]: r["launches"] for r in _lib.profile_report(DEV)}
        used = sum(fwd_rows.values())
        _lib.profile_enable(DEV, "gemm_tc_nt")
        y.backward(to_t(dy))
        torch.cuda.synchronize()
        bwd_tags = {r["tag"] for r in _lib.profile_report(DEV)}
        _lib.profile_enable(DEV, None)
        assert _lib.async_error(DEV) == 0
        if H >= 128 and B % 4 == 0:
            assert {"gemm_tc/lstm_dWih", "gemm_tc/lstm_dWhh", "gemm_tc/lstm_dx"} <= bwd_tags, bwd_tags
        # H = 64 keeps its persistent FMA recurrent kernel: only layer 1's input projection qualifies there; H = 128 has the
        # persistent tcgen05 kernel (one launch per layer); larger H: one fused tcgen05 launch per timestep and layer
        if H == 128:
            # persistent kernel, one launch per layer, on the chunked gate buffer written by the input-projection GEMM / the
            # layer-0 projection kernels - in the grad-carrying pass (chunked stash, read by the chunked BPTT kernel) ...
            assert fwd_rows.get("lstm128_tc_fwd_kernel") == L and "gemm_tc_lstm_fwd_kernel" not in fwd_rows, fwd_rows
        else:
            assert used >= (1 if H <= 64 else T), f"the tcgen05 GEMM / fused step kernels took only {used} launches"
        if H == 128:
            # ... and in the no-grad passes (sampling, the critic phase's generations)
            _lib.profile_enable(DEV, "tc_")
            with torch.no_grad():
                y_ng = G(to_t(proto), to_t(z))
            ng_rows = {r["tag"]: r["launches"] for r in _lib.profile_report(DEV)}
            _lib.profile_enable(DEV, None)
            assert ng_rows.get("lstm128_tc_fwd_kernel") == L and "gemm_tc_lstm_fwd_kernel" not in ng_rows, ng_rows
            assert max_abs_rel(to_np(y_ng), y_ref) <= 3e-3
        e_fwd = max_abs_rel(to_np(y), y_ref)
        worst = max(rel_l2(v, g_ref[k]) for k, v in grads_of(G).items())
        print(f"H={H} B={B} T={T}: {used} tcgen05 GEMM launches, fwd {e_fwd:.2e}, grads {worst:.2e}")
        assert e_fwd <= 3e-3 and worst <= 3e-3 and rel_l2(to_np(zt.grad), dz_ref) <= 3e-3, (e_fwd, worst)
    finally:
        wgg.set_math_mode("fp32")


def test_train_batch_h128_t256():
    """One whole training batch of the scaled model (H = 128, T = 256, 4 layers, TemporalDiscriminator on 256-point
    gestures) against the fp64 CPU restatement, fp32 mode: all 11 losses."""
    ocfg = O.ModelCfg(seq_length=256, gen_hidden_dim=128)
    B = 4
    wgg.seed_everything(7)
    tr = wgg.WordGestureGANTrainer(model_cfg(ocfg), wgg.TrainingConfig(), DEV)
    tp = torch_port.TorchPortTrainer(seed=0, cfg=ocfg, tc=O.TrainCfg(), dtype=torch.float64)
    tp.load_state({m: state_of(getattr(tr, ATTR[m])) for m in ("G", "E", "D1", "D2")})
    for m in ("G", "E", "D1", "D2"):
        getattr(tr, ATTR[m]).train()
    rng = np.random.default_rng(3)
    f32 = lambda a: a.astype(np.float32).astype(np.float64)
    real = f32(rng.uniform(-1, 1, (B, 256, 3)))
    proto = f32(rng.uniform(-1, 1, (B, 256, 3)))
    noise = [f32(rng.standard_normal((B, 32))) for _ in range(13)]
    ref = tp.train_batch(torch.from_numpy(real), torch.from_numpy(proto), noise=noise)
    out = wgg.train_batch(tr, to_t(real), to_t(proto), 1.0, [to_t(n) for n in noise])
    for k in LOSS_KEYS:
        assert abs(out[k].item() - ref[k]) <= 2e-3 * max(abs(ref[k]), 1e-2), (k, out[k].item(), ref[k])


def test_train_batch_h128_t256_tf32_lr0():
    """The same batch in mode tf32 - fused tcgen05 step kernels, split-K tcgen05 weight-gradient GEMMs and the tcgen05 conv
    kernels on 256-point gestures (two 128-row tiles per gesture) - at learning rate 0 (the first Adam steps are sign-like
    and would amplify TF32-level differences between critic iterations): all 11 losses at the TF32 tolerance
    of tests/test_gpu_parity_tc.py (5e-3)."""
    from wgg_b200 import _lib
    ocfg = O.ModelCfg(seq_length=256, gen_hidden_dim=128, gen_num_layers=2)
    B = 132
    wgg.set_math_mode("tf32")
    try:
        wgg.seed_everything(11)
        tr = wgg.WordGestureGANTrainer(model_cfg(ocfg), wgg.TrainingConfig(learning_rate=0.0), DEV)
        tp = torch_port.TorchPortTrainer(seed=0, cfg=ocfg, tc=O.TrainCfg(learning_rate=0.0), dtype=torch.float64)
        tp.load_state({m: state_of(getattr(tr, ATTR[m])) for m in ("G", "E", "D1", "D2")})
        for m in ("G", "E", "D1", "D2"):
            getattr(tr, ATTR[m]).train()
        rng = np.random.default_rng(5)
        f32 = lambda a: a.astype(np.float32).astype(np.float64)
        real = f32(rng.uniform(-1, 1, (B, 256, 3)))
        proto = f32(rng.uniform(-1, 1, (B, 256, 3)))
        noise = [f32(rng.standard_normal((B, 32))) for _ in range(13)]
        ref = tp.train_batch(torch.from_numpy(real), torch.from_numpy(proto), noise=noise)
        out = wgg.train_batch(tr, to_t(real), to_t(proto), 1.0, [to_t(n) for n in noise])
        torch.cuda.synchronize()
        assert _lib.async_error(DEV) == 0
        for k in LOSS_KEYS:
            assert abs(out[k].item() - ref[k]) <= 5e-3 * max(abs(ref[k]), 1e-2), (k, out[k].item(), ref[k])
    finally:
        wgg.set_math_mode("fp32")
