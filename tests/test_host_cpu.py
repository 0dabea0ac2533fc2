"""CPU-only checks of the host layer: the C-ABI library loads and exports every symbol include/wgg.h declares,
flat-parameter layouts agree between Python and C, construction reproduces the reference's state-dict contract,
and compute entry points refuse to run without a CUDA device (no silent fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import wgg_b200 as wgg
from golden_util import Golden, summarise
from wgg_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "wgg.h")).read()
    return sorted(set(re.findall(r"WGG_API\s+[\w\s\*]+?\b(wgg_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    syms = header_symbols()
    assert len(syms) >= 39
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(handle, s), f"{s} declared in include/wgg.h but not exported"
    assert sorted(_lib.EXPORTED_SYMBOLS) == syms, "ctypes signature table and header disagree"
    assert _lib.lib().wgg_abi_version() == 1


@pytest.mark.parametrize("kw", [dict(), dict(use_temporal_disc=False), dict(prototype_has_time=True, gen_hidden_dim=32),
                                dict(seq_length=16, latent_dim=4, gen_hidden_dim=8, gen_num_layers=2,
                                     enc_hidden_dims=(24, 12, 8, 6), disc_hidden_dims=(20, 12, 8, 6))])
def test_flat_layout_sizes_agree_with_c(kw):
    mc = wgg.ModelConfig(**kw)
    lib = _lib.lib()
    cfg = _lib.c_cfg(mc)
    G, E = wgg.Generator(mc), wgg.VariationalEncoder(mc)
    D = (wgg.TemporalDiscriminator if mc.use_temporal_disc else wgg.Discriminator)(mc)
    assert lib.wgg_generator_param_floats(cfg) == G.flat_params().numel()
    assert lib.wgg_encoder_param_floats(cfg) == E.flat_params().numel()
    assert lib.wgg_disc_param_floats(cfg) == D.flat_params().numel()
    assert lib.wgg_disc_uv_floats(cfg) == D.flat_buffers().numel()
    # parameters are views of the flat buffer, in named_parameters order
    off = 0
    flat = G.flat_params()
    for p in G.parameters():
        assert p.data_ptr() == flat.data_ptr() + 4 * off
        off += p.numel()
    B = 7
    nf = lib.wgg_disc_num_features(cfg)
    widths = [lib.wgg_disc_feature_width(cfg, k) for k in range(nf)]
    if mc.use_temporal_disc:
        assert widths == [mc.seq_length * 64, mc.seq_length * 64, mc.seq_length * 32, 128, 64]
    else:
        assert widths == list(mc.disc_hidden_dims)
    assert lib.wgg_disc_stash_floats(cfg, B) >= B * sum(widths)
    # large enough for the FMA-path layout (x0 + per layer hseq/gates/c); the tcgen05 layout may need more
    assert lib.wgg_generator_stash_floats(cfg, B) >= mc.seq_length * B * (
        (3 if mc.prototype_has_time else 2) + mc.latent_dim + 12 * mc.gen_hidden_dim * mc.gen_num_layers)


@pytest.mark.parametrize("H,T,L", [(128, 256, 4), (96, 64, 3), (512, 256, 2)])
def test_scaled_regime_buffer_sizes(H, T, L):
    """Scaled regime (BASELINE configs[3]; csrc/lstm.cu stash_floats / wgg_generator_workspace_floats): the stash holds x0 and
    hseq at the true batch, the gate / c buffers padded to whole 128-gesture tiles (chunked order of the tcgen05 recurrence
    kernels) and the per-gesture latent term of the layer-0 projection; the sizes are host-side functions of (config, batch)
    alone, grow with the batch and cover the row-major layout as well."""
    mc = wgg.ModelConfig(gen_hidden_dim=H, seq_length=T, gen_num_layers=L)
    lib, cfg = _lib.lib(), _lib.c_cfg(mc)
    I0 = (3 if mc.prototype_has_time else 2) + mc.latent_dim
    prev = (0, 0, 0)
    for B in (1, 5, 128, 129, 1024):
        Bp = (B + 127) // 128 * 128
        stash = lib.wgg_generator_stash_floats(cfg, B)
        assert stash >= T * B * (I0 + 2 * H * L) + T * Bp * 10 * H * L + 8 * Bp * H
        assert stash >= T * B * (I0 + 12 * H * L)                      # the row-major layout fits too
        fwd, bwd = lib.wgg_generator_workspace_floats(cfg, B, 0), lib.wgg_generator_workspace_floats(cfg, B, 1)
        assert fwd >= T * B * (I0 + 4 * H) + T * Bp * 8 * H + 8 * Bp * H  # x0, two hseq, padded gate buffer, latent term
        assert bwd >= T * B * (3 + 2 * 2 * H) + T * B * (8 * H + 4 * H + 8 * H)  # dpre/dh/dx + daT, hT x2, da row-major copy
        assert (stash, fwd, bwd) > prev or B == 1
        assert stash >= prev[0] and fwd >= prev[1] and bwd >= prev[2]
        prev = (stash, fwd, bwd)


def test_state_dict_contract_default():
    """Keys / shapes / order listed in SURVEY.md section 8b (measured on the reference)."""
    G = wgg.Generator()
    keys = list(G.state_dict())
    assert keys[:4] == ["lstm.weight_ih_l0", "lstm.weight_hh_l0", "lstm.bias_ih_l0", "lstm.bias_hh_l0"]
    assert keys[4] == "lstm.weight_ih_l0_reverse" and keys[-2:] == ["output_layer.weight", "output_layer.bias"]
    assert tuple(G.state_dict()["lstm.weight_ih_l0"].shape) == (192, 34)
    assert tuple(G.state_dict()["lstm.weight_ih_l1"].shape) == (192, 96)
    E = wgg.VariationalEncoder()
    assert list(E.state_dict()) == ["encoder.0.weight", "encoder.0.bias", "encoder.2.weight", "encoder.2.bias",
                                    "encoder.4.weight", "encoder.4.bias", "encoder.6.weight", "encoder.6.bias",
                                    "fc_mu.weight", "fc_mu.bias", "fc_log_var.weight", "fc_log_var.bias"]
    D = wgg.TemporalDiscriminator()
    assert list(D.state_dict())[:4] == ["temporal_conv.0.bias", "temporal_conv.0.weight_orig",
                                        "temporal_conv.0.weight_u", "temporal_conv.0.weight_v"]
    assert [k for k, _ in D.named_parameters()][-2:] == ["output_layer.bias", "output_layer.weight_orig"]
    assert tuple(D.state_dict()["temporal_conv.2.weight_v"].shape) == (320,)
    M = wgg.Discriminator()
    assert list(M.state_dict())[:4] == ["layers.0.bias", "layers.0.weight_orig", "layers.0.weight_u", "layers.0.weight_v"]
    assert sum(p.numel() for p in G.parameters()) == 200739
    assert sum(p.numel() for p in E.parameters()) == 100784
    assert sum(p.numel() for p in D.parameters()) == 68961
    assert sum(p.numel() for p in M.parameters()) == 98305


def test_seed42_init_matches_reference_fixture():
    """seed_everything(42) + trainer construction reproduces the reference's initial weights bit for bit
    (fixture written by oracle/make_golden.py from the reference's own constructor)."""
    z = np.load(os.path.join(ROOT, "tests", "golden", "init_seed42.npz"))
    wgg.seed_everything(42)
    tr = wgg.WordGestureGANTrainer(device="cpu")
    for n, attr in (("G", "generator"), ("E", "encoder"), ("D1", "discriminator_1"), ("D2", "discriminator_2")):
        sd = getattr(tr, attr).state_dict()
        assert list(sd) == [str(k) for k in z[f"order/{n}"]]
        for k, v in sd.items():
            assert np.array_equal(summarise(v.numpy()), z[f"{n}/{k}"]), (n, k)


def test_load_state_dict_keeps_flat_aliasing_and_deepcopy():
    import copy
    G = wgg.Generator()
    sd = {k: torch.randn_like(v) for k, v in G.state_dict().items()}
    G.load_state_dict(sd)
    assert G._is_flat()
    assert torch.equal(G.flat_params()[:192 * 34].view(192, 34), sd["lstm.weight_ih_l0"])
    G2 = copy.deepcopy(G)
    assert G2._is_flat() and G2.flat_params().data_ptr() != G.flat_params().data_ptr()
    assert torch.equal(G2.flat_params(), G.flat_params())


def test_no_cpu_fallback():
    G = wgg.Generator()
    with pytest.raises(_lib.WggError):
        G(torch.zeros(2, 128, 3), torch.zeros(2, 32))
    with pytest.raises(_lib.WggError):
        wgg.VariationalEncoder()(torch.zeros(2, 128, 3))
    with pytest.raises(_lib.WggError):
        wgg.TemporalDiscriminator()(torch.zeros(2, 128, 3))


def test_dropin_import_paths():
    from wgg_b200 import dropin
    dropin.install_as_src(force=True)
    import importlib
    m = importlib.import_module("src.gan.models")
    assert m.Generator is wgg.Generator and m.TemporalDiscriminator is wgg.TemporalDiscriminator
    t = importlib.import_module("src.gan.trainer")
    assert t.WordGestureGANTrainer is wgg.WordGestureGANTrainer
    u = importlib.import_module("src.shared.utils")
    assert u.train_epoch_with_grad_clip is wgg.train_epoch_with_grad_clip and u.seed_everything is wgg.seed_everything
    c = importlib.import_module("src.shared.config")
    assert c.ModelConfig().gen_hidden_dim == 48 and c.TrainingConfig().n_critic == 5
    from src.gan.losses import WassersteinLoss  # noqa
    for k in [k for k in list(__import__("sys").modules) if k == "src" or k.startswith("src.")]:
        del __import__("sys").modules[k]


def test_optimizer_is_torch_optimizer_and_scheduler_compatible():
    E = wgg.VariationalEncoder()
    opt = wgg.FusedClipAdam(E, lr=2e-4, betas=(0.5, 0.999))
    assert isinstance(opt, torch.optim.Optimizer)
    sd = opt.state_dict()
    assert sd["state"] == {} and sd["param_groups"][0]["params"] == list(range(12))
    assert sd["param_groups"][0]["betas"] == (0.5, 0.999)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=200, eta_min=1e-5)
    assert sched.get_last_lr()[0] == pytest.approx(2e-4)
    # a reference-format Adam state dict (torch.optim.Adam) loads
    ref_opt = torch.optim.Adam(wgg.VariationalEncoder().parameters(), lr=2e-4, betas=(0.5, 0.999))
    for p in ref_opt.param_groups[0]["params"]:
        p.grad = torch.ones_like(p)
    ref_opt.step()
    opt.load_state_dict(ref_opt.state_dict())
    assert opt._step == 1
    assert torch.allclose(opt._m, torch.full_like(opt._m, 0.5))


def test_ctypes_signatures_match_header_arity_and_types():
    """Every ctypes signature has the header's parameter count and pointer/integer/float kinds."""
    from ctypes import c_char_p, c_float, c_int, c_int32, c_int64, c_void_p
    text = open(os.path.join(ROOT, "include", "wgg.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = re.findall(r"WGG_API\s+([\w\s\*]+?)\b(wgg_\w+)\s*\(([^)]*)\)\s*;", text)
    assert len(protos) == len(_lib.EXPORTED_SYMBOLS)
    for ret, name, params in protos:
        params = params.strip()
        plist = [] if params in ("", "void") else [p.strip() for p in params.split(",")]
        res, args = _lib._SIG[name]
        assert len(args) == len(plist), f"{name}: header has {len(plist)} parameters, ctypes table has {len(args)}"
        for decl, a in zip(plist, args):
            if "*" in decl:
                assert a not in (c_int, c_int32, c_int64, c_float), (name, decl, a)
            elif decl.startswith("int64_t"):
                assert a is c_int64, (name, decl, a)
            elif decl.startswith("float"):
                assert a is c_float, (name, decl, a)
            elif decl.startswith(("int32_t", "int ")):
                assert a in (c_int, c_int32), (name, decl, a)
            else:
                raise AssertionError(f"{name}: unhandled parameter declaration {decl!r}")


def test_device_resident_loader_matches_dataloader_contract():
    """the resident loader yields the reference's batch dicts (data.py:157-164, 526-533) - every
    sample exactly once per epoch, gesture/prototype rows stay paired, seeded shuffles reproduce, drop_last / ragged
    last batch follow torch.utils.data.DataLoader."""
    import torch
    import wgg_b200 as wgg
    n, T = 23, 16
    g = torch.arange(n, dtype=torch.float32).view(n, 1, 1).expand(n, T, 3).contiguous()
    p = -g
    words = [f"w{i}" for i in range(n)]
    ld = wgg.DeviceResidentLoader(g, p, batch_size=8, shuffle=True, generator=torch.Generator().manual_seed(3), words=words)
    assert len(ld) == 3
    seen = []
    for b in ld:
        assert set(b) == {"gesture", "prototype", "word"}
        assert b["gesture"].shape[1:] == (T, 3) and torch.equal(b["gesture"], -b["prototype"])
        ids = b["gesture"][:, 0, 0].long().tolist()
        assert b["word"] == [f"w{i}" for i in ids]
        seen += ids
    assert sorted(seen) == list(range(n)) and seen != list(range(n))
    again = [i for b in wgg.DeviceResidentLoader(g, p, 8, True, generator=torch.Generator().manual_seed(3)) for i in b["gesture"][:, 0, 0].long().tolist()]
    assert again == seen
    ld2 = wgg.DeviceResidentLoader(g, p, batch_size=8, shuffle=False, drop_last=True)
    sizes = [b["gesture"].size(0) for b in ld2]
    assert sizes == [8, 8] and len(ld2) == 2
    first = next(iter(ld2))["gesture"][:, 0, 0].tolist()
    assert first == list(range(8))
    import pytest
    with pytest.raises(ValueError):
        wgg.DeviceResidentLoader(g, p[:, :8], 4)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm of the measurement contract) prints ONE JSON line with the keys the
    driver reads; it runs entirely on the host (tiny bounded sample here)."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--ref-batch", "8"], capture_output=True, text=True, timeout=600, check=True)
    lines = [l for l in out.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "gestures/s" and d["value"] > 0
    # "reference": the reference's own files were reachable (/root/reference or the vendored oracle/_ref) and were
    # what ran; "port": oracle/torch_port.py stood in for them
    from oracle.ref_loader import reference_available
    assert d["cpu_baseline"]["kind"] == ("reference" if reference_available() else "port")
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and "workload" in d["config"]
