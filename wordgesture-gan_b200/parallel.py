"""Data-parallel plumbing: one process per GPU, shard the gesture minibatch, all-reduce flat gradients.

The reference is single-GPU; this is the new data-parallel wrapper the north star asks for
(SURVEY.md section 8e).  Every loss of the step is a batch mean and no layer mixes samples, so averaging the
per-rank flat gradients of equal shards reproduces the global-batch gradient; spectral-norm power iterations
depend on weights only, so replicas stay identical with no further communication.  The collective is NCCL
(over NVLink 5 / NVSwitch) on the flat gradient buffer, issued BEFORE clipping (the clip is on the global
gradient).  Noise is drawn for the GLOBAL batch with the shared seed on every rank and sliced, so an N-rank
run consumes the same random numbers as a 1-rank run on the same global batch.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous, equal shards: rank r owns [r*n/W, (r+1)*n/W).  n must divide evenly (batch means!)."""
    if n % world_size != 0:
        raise ValueError(f"global batch {n} is not divisible by world size {world_size}: unequal shards would "
                         "bias the averaged gradient")
    per = n // world_size
    return rank * per, (rank + 1) * per


def shard_batch(t: torch.Tensor, rank: int, world_size: int) -> torch.Tensor:
    lo, hi = shard_bounds(t.shape[0], rank, world_size)
    return t[lo:hi]


def draw_global_noise(n_draws: int, global_batch: int, latent_dim: int, device, generator=None) -> List[torch.Tensor]:
    """The step's (B, Z) normal draws for the GLOBAL batch, as individual torch.randn calls in the
    reference's consumption order; identical on every rank when the seeds agree."""
    return [torch.randn(global_batch, latent_dim, device=device, generator=generator) for _ in range(n_draws)]


def randn_rank_rows(per_rank_batch: int, latent_dim: int, rank: int, world_size: int, device, generator=None
                    ) -> torch.Tensor:
    """One (B, Z) normal draw of the step as the product path makes it under data parallelism: EVERY rank draws the
    GLOBAL (world * B, Z) tensor - same seed, same generator offset on all ranks, so the same numbers - and keeps
    its own rows.  The concatenation over ranks is exactly the draw a single process would make for the global
    batch (utils.py:71, trainer.py:105, models.py:85), and the generators stay in lock step."""
    if world_size == 1:
        return torch.randn(per_rank_batch, latent_dim, device=device, generator=generator)
    full = torch.randn(world_size * per_rank_batch, latent_dim, device=device, generator=generator)
    return full[rank * per_rank_batch:(rank + 1) * per_rank_batch].contiguous()


def allreduce_mean_(flat: torch.Tensor, group=None, world_size: Optional[int] = None) -> torch.Tensor:
    """In-place mean all-reduce of a flat gradient bucket."""
    if world_size is None:
        world_size = dist.get_world_size(group)
    if world_size == 1:
        return flat
    backend = dist.get_backend(group)
    if backend == "nccl":
        dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=group)
    else:  # gloo (CPU tests) has no AVG
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.div_(world_size)
    return flat


class _DeviceMemory:
    """Raw device memory exposed to torch through __cuda_array_interface__ (torch.as_tensor aliases it, no copy)."""

    def __init__(self, ptr: int, nfloats: int):
        self.__cuda_array_interface__ = {"shape": (nfloats,), "typestr": "<f4", "data": (ptr, False), "version": 3}


class P2PBucket:
    """A module's flat gradient bucket placed in memory that every rank of the node has mapped (CUDA IPC), reduced by
    the one-shot peer-memory kernel ``wgg_p2p_allreduce_avg`` (csrc/optim.cu) instead of an NCCL collective: the
    buckets of this model are 0.3 - 1.2 MB, so the exchange is pure latency; reading the peers' buckets directly over
    NVLink and summing them in rank order takes one small launch and keeps replicas bit-identical.

    Setup (once): each rank allocates [bucket | flag block] with ``wgg_p2p_alloc`` (cudaMalloc + IPC handle),
    all-gathers the 64-byte handles over the process group and maps the peers' allocations with ``wgg_p2p_open``; the
    module's persistent gradient buffer (FlatModule._gflat) is redirected INTO the shared allocation, so the backward
    kernels write where the peers read - no staging copy."""

    def __init__(self, module, group, rank: int, world: int):
        import ctypes
        from . import _lib
        lib = _lib.lib()
        flat = module.flat_params()
        dev = flat.device
        c = _lib.ctx(dev)
        self.n = flat.numel()
        self.npad = (self.n + 3) // 4 * 4
        words = int(lib.wgg_p2p_flag_words())
        total = self.npad + words
        self.rank, self.world = rank, world
        # Every rank takes part in every collective below whatever happens locally, and failure is decided
        # collectively - a rank that fell back to NCCL while its peers spin on flags would hang the job.
        own = ctypes.c_void_p()
        handle = ctypes.create_string_buffer(64)
        rc = lib.wgg_p2p_alloc(c, 4 * total, ctypes.byref(own), handle)
        meta = bytes(handle.raw) if rc == 0 else f"wgg_p2p_alloc failed: {lib.wgg_last_error(c).decode()}"
        gathered = [None] * world
        dist.all_gather_object(gathered, meta, group=group)
        err = next((m for m in gathered if isinstance(m, str)), None)
        ptrs = []
        if err is None:
            for r, h in enumerate(gathered):
                if r == rank:
                    ptrs.append(own.value)
                    continue
                p = ctypes.c_void_p()
                if lib.wgg_p2p_open(c, h, ctypes.byref(p)) != 0:
                    err = f"wgg_p2p_open(rank {r}) failed: {lib.wgg_last_error(c).decode()}"
                    break
                ptrs.append(p.value)
        torch.cuda.set_device(dev)
        ok = torch.tensor([0 if err else 1], dtype=torch.int32, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok.item()) == 0:
            raise RuntimeError(f"peer mapping failed on at least one rank ({err or 'on a peer'})")
        self._mem = _DeviceMemory(own.value, total)            # keeps the interface object alive
        self.shared = torch.as_tensor(self._mem, device=dev)   # aliases the cudaMalloc'ed bucket + flag block
        assert self.shared.data_ptr() == own.value and self.shared.numel() == total
        self.peer_ptrs = ptrs
        self.grad_ptrs = torch.tensor(ptrs, dtype=torch.int64, device=dev)
        self.flag_ptrs = torch.tensor([p + 4 * self.npad for p in ptrs], dtype=torch.int64, device=dev)
        self.avg = torch.zeros(self.npad, dtype=torch.float32, device=dev)
        self.state = torch.zeros(4, dtype=torch.int32, device=dev)
        self.module = module
        # the module's gradient buffer now lives in the shared allocation (padding floats stay zero)
        module._gflat = self.shared[:self.n]
        module.grad_buffer()
        torch.cuda.synchronize(dev)
        dist.barrier(group=group)   # every rank has mapped every bucket before the first reduce

    def allreduce_mean_(self):
        """In-place mean of the module's gradient bucket over all ranks (asynchronous, graph-capturable)."""
        from . import _lib
        dev = self.shared.device
        c = _lib.ctx(dev)
        _lib.check(_lib.lib().wgg_p2p_allreduce_avg(c, self.grad_ptrs.data_ptr(), self.flag_ptrs.data_ptr(), self.rank,
                                                    self.world, self.npad, self.avg.data_ptr(), self.shared.data_ptr(),
                                                    self.state.data_ptr(), _lib.stream(dev)), c)


def p2p_enabled() -> bool:
    """The peer-memory exchange is opt-in (WGG_P2P=1).  Measured on 8 x B200 (profiles/r02_bench_B4096_tf32_dp8_*.json):
    32.48 ms per step with it, 32.29 ms with NCCL's all-reduce - at these bucket sizes NCCL's low-latency protocol is
    as fast as two flag round trips over NVLink, so NCCL stays the default."""
    import os
    return os.environ.get("WGG_P2P", "0") == "1"


class DataParallelGAN:
    """Attaches a process group to a trainer's four fused optimisers: each ``step`` first mean-all-reduces the
    module's flat gradient bucket.  Also broadcasts rank 0's state so replicas start identical."""

    def __init__(self, trainer, group=None):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.trainer = trainer
        self.group = group
        self.rank = dist.get_rank(group)
        self.world_size = dist.get_world_size(group)
        for opt in (trainer.optimizer_G, trainer.optimizer_E, trainer.optimizer_D1, trainer.optimizer_D2):
            opt.process_group = group if group is not None else dist.group.WORLD
            opt.world_size = self.world_size
        # the D2 chain of the critic phase runs on its own stream next to the D1 chain (train_step.train_batch):
        # give its all-reduces their own communicator so that collectives of the two chains never share one
        ranks = dist.get_process_group_ranks(group) if group is not None else list(range(dist.get_world_size()))
        self.group_d2 = dist.new_group(ranks=ranks)
        trainer.optimizer_D2.process_group = self.group_d2
        trainer._dp = self  # train_step.train_batch draws the step's noise for the global batch and slices by rank
        self.sync_state()
        # Gradient exchange: NCCL all-reduce by default; the one-shot peer-memory reduce (P2PBucket) on request
        # (WGG_P2P=1) when every rank sits on a CUDA device of this node (falls back if the IPC mapping fails).
        self.p2p = False
        dev = trainer.generator.flat_params().device
        if dev.type == "cuda" and self.world_size > 1 and p2p_enabled():
            try:
                for opt in (trainer.optimizer_G, trainer.optimizer_E, trainer.optimizer_D1, trainer.optimizer_D2):
                    opt.p2p = P2PBucket(opt.module, self.group, self.rank, self.world_size)
                self.p2p = True
            except RuntimeError as ex:   # raised on EVERY rank (collective decision): keep the NCCL path; say why
                import warnings
                warnings.warn(f"peer-memory gradient exchange unavailable ({ex!r}); using NCCL all-reduce")
                for opt in (trainer.optimizer_G, trainer.optimizer_E, trainer.optimizer_D1, trainer.optimizer_D2):
                    opt.p2p = None

    def sync_state(self):
        for mod in (self.trainer.generator, self.trainer.encoder, self.trainer.discriminator_1,
                    self.trainer.discriminator_2):
            dist.broadcast(mod.flat_params(), src=0, group=self.group)
            fb = mod.flat_buffers()
            if fb is not None:
                dist.broadcast(fb, src=0, group=self.group)

    def shard(self, t: torch.Tensor) -> torch.Tensor:
        return shard_batch(t, self.rank, self.world_size)
