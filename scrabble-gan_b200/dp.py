"""Data parallelism: one process per GPU, torch.distributed (NCCL over NVLink/NVSwitch) as plumbing.

The train step is pure data parallel over the batch with exactly three exchange points:
  (1) batch-norm statistics of the generator: raw sums [2C] forward, [2C] backward per BN layer      (tiny, latency bound)
  (2) gradient balancing / loss statistics: 16 doubles                                               (tiny)
  (3) parameter gradients: ONE sum-all-reduce per network over its flat gradient bucket (SUM, not mean: the
      reference differentiates sums over the batch)
All three go through Runtime.allreduce_.  All ranks must use the same (L_real, L_fake) per step -- one length bucket
per step is already the reference's own rule (data_utils.py:64,386)."""
from __future__ import annotations

import os

import torch


def init_data_parallel(rt, backend: str = None) -> None:
    """Initialise torch.distributed from the torchrun environment and attach it to the runtime."""
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        rt.world_size, rt.rank = 1, 0
        return
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        kw = {}
        if backend == "nccl":
            kw["device_id"] = rt.device
        dist.init_process_group(backend=backend, **kw)
    rt.world_size = dist.get_world_size()
    rt.rank = dist.get_rank()
    rt.process_group = None


def broadcast_parameters(rt, models) -> None:
    """Make every replica start from rank 0's weights (and BN moving statistics)."""
    if rt.world_size <= 1:
        return
    import torch.distributed as dist
    for m in models:
        dist.broadcast(m.store.w, src=0)
        dist.broadcast(m.store.s, src=0)
        m.store.version += 1


def shard_batch(global_batch: int, world_size: int, rank: int):
    """Even contiguous shards of the global batch; returns (start, stop)."""
    assert global_batch % world_size == 0, "global batch must divide evenly across replicas"
    per = global_batch // world_size
    return rank * per, (rank + 1) * per


def length_schedule(step: int, seed: int = 1234, max_len: int = 10, fixed=None):
    """(L_real, L_fake) for a step, identical on every rank (derived from the shared seed only)."""
    if fixed is not None:
        return fixed, fixed
    import random
    r = random.Random(seed * 1000003 + step)
    return r.randint(1, max_len), r.randint(1, max_len)
