"""CPU, world_size 2, gloo: the host-side logic of the data-parallel path - sharding,
global-noise slicing, replica synchronisation and the flat-bucket mean all-reduce that precedes clipping."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import wgg_b200 as wgg
        from wgg_b200 import parallel

        # 1. equal contiguous shards; unequal shards are refused
        assert parallel.shard_bounds(8, rank, world) == (rank * 4, rank * 4 + 4)
        try:
            parallel.shard_bounds(7, rank, world)
            raise AssertionError("uneven shard accepted")
        except ValueError:
            pass

        # 2. global noise: same tensors on every rank (shared seed), shards are the matching rows
        torch.manual_seed(123)
        noise = parallel.draw_global_noise(13, 8, 32, "cpu")
        gathered = [torch.zeros_like(noise[5]) for _ in range(world)]
        dist.all_gather(gathered, noise[5])
        assert all(torch.equal(g, noise[5]) for g in gathered)
        mine = parallel.shard_batch(noise[5], rank, world)
        assert torch.equal(mine, noise[5][rank * 4:(rank + 1) * 4])

        # 3. replicas start from rank 0's state
        wgg.seed_everything(100 + rank)  # deliberately different initial weights per rank
        tr = wgg.WordGestureGANTrainer(wgg.ModelConfig(seq_length=16, latent_dim=4, gen_hidden_dim=8, gen_num_layers=2,
                                                       enc_hidden_dims=(24, 12, 8, 6)), wgg.TrainingConfig(), "cpu")
        dp = parallel.DataParallelGAN(tr)
        for mod in (tr.generator, tr.encoder, tr.discriminator_1, tr.discriminator_2):
            flat = mod.flat_params()
            ref = flat.clone()
            dist.broadcast(ref, src=0)
            assert torch.equal(flat, ref)
            fb = mod.flat_buffers()
            if fb is not None:
                refb = fb.clone()
                dist.broadcast(refb, src=0)
                assert torch.equal(fb, refb)
        assert tr.optimizer_G.world_size == world and tr.optimizer_D1.process_group is not None

        # 4. mean all-reduce of per-shard flat gradients == global-batch gradient (every loss is a batch mean)
        torch.manual_seed(7)
        W = torch.randn(5, 3, dtype=torch.float64, requires_grad=True)
        X = torch.randn(8, 3, dtype=torch.float64)
        loss_global = (X @ W.t()).tanh().mean(dim=1).mean()  # mean over the batch of a per-sample quantity
        g_global, = torch.autograd.grad(loss_global, W)
        Xs = dp.shard(X)
        loss_local = (Xs @ W.t()).tanh().mean(dim=1).mean()
        g_local, = torch.autograd.grad(loss_local, W)
        flat = g_local.reshape(-1).clone()
        parallel.allreduce_mean_(flat, None, world)
        assert torch.allclose(flat, g_global.reshape(-1), rtol=1e-12, atol=1e-15)

        # 5. the optimiser's flat gradient bucket is one contiguous tensor in parameter order
        opt = tr.optimizer_E
        for i, p in enumerate(opt.param_groups[0]["params"]):
            p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
        bucket = opt.flat_grad()
        assert bucket.numel() == sum(p.numel() for p in tr.encoder.parameters())
        parallel.allreduce_mean_(bucket, None, world)
        first = opt.param_groups[0]["params"][0]
        assert torch.allclose(first.grad, torch.full_like(first, 1.5))  # .grad aliases the reduced bucket

        # 6. the PRODUCT path's noise under data parallelism (train_step.train_batch / trainer._randn): every rank
        # seeds alike, draws the global tensor and keeps its rows - ranks get DIFFERENT rows, their concatenation is
        # the single-process draw for the global batch, and the generators stay in lock step over successive draws
        from wgg_b200 import train_step
        assert train_step.dp_slice(tr) == (rank, world)
        wgg.seed_everything(42)
        draws = [tr._randn(4) for _ in range(3)]
        wgg.seed_everything(42)
        ref = [torch.randn(world * 4, tr.model_config.latent_dim) for _ in range(3)]
        for d, r in zip(draws, ref):
            assert torch.equal(d, r[rank * 4:(rank + 1) * 4])
            both = [torch.zeros_like(d) for _ in range(world)]
            dist.all_gather(both, d)
            assert torch.equal(torch.cat(both, 0), r)
            assert not torch.equal(both[0], both[1])
        ret[rank] = "ok"
    except Exception as ex:  # pragma: no cover
        import traceback
        ret[rank] = "".join(traceback.format_exception(ex))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_data_parallel_host_logic_world2():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    mgr = ctx.Manager()
    ret = mgr.dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(240)
    for r in range(world):
        assert ret.get(r) == "ok", ret.get(r)
