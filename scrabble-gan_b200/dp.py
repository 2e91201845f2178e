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
    else:
        backend = str(dist.get_backend())           # a process group made by the caller: ask it what it is
    rt.world_size = dist.get_world_size()
    rt.rank = dist.get_rank()
    rt.process_group = None
    rt.peer = None
    rt.peer_comm = None
    if "nccl" in backend:
        # every rank must take the SAME path (a rank on NCCL while its peers spin on peer flags would hang both): the
        # opt-out and the outcome of the set-up are agreed on with a MIN all-reduce
        want = 0 if os.environ.get("SGAN_NO_PEER", "0") == "1" else 1
        flag = torch.tensor([want], device=rt.device, dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok = int(flag.item()) == 1 and init_peer_exchange(rt)
        flag = torch.tensor([1 if ok else 0], device=rt.device, dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) != 1:
            rt.peer = rt.peer_comm = None


class PeerExchange:
    """Symmetric-memory buffer + peer pointer table for the one-shot NVLink exchanges of libsgan (csrc/peer.cu)."""

    def __init__(self, buf, handle, ptrs, world, rank):
        import ctypes as C
        self.buf, self.handle = buf, handle            # keep the allocation and the rendezvous handle alive
        self.world, self.rank = world, rank
        self.ptrs = (C.c_ulonglong * world)(*[int(p) for p in ptrs])


def init_peer_exchange(rt) -> bool:
    """Allocate the small exchange buffer in symmetric memory (torch.distributed._symmetric_memory: plumbing only) and map
    every peer's copy.  On any failure the runtime keeps using NCCL for the small exchanges (rt.peer stays None)."""
    import sys
    import torch.distributed as dist
    from . import _abi
    rt.peer = None
    try:
        import torch.distributed._symmetric_memory as symm_mem
        nbytes = int(_abi.load().sg_peer_buffer_bytes())
        group = dist.group.WORLD
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            try:                                   # required by older torch releases, a deprecated no-op in newer ones
                symm_mem.enable_symm_mem_for_group(group.group_name)
            except Exception:
                pass
        made = []
        for _ in range(2):       # [0]: the compute stream's small exchanges; [1]: barrier flags of the communication stream
            buf = symm_mem.empty(nbytes, dtype=torch.uint8, device=rt.device)
            hdl = symm_mem.rendezvous(buf, group.group_name)
            buf.zero_()
            torch.cuda.synchronize(rt.device)
            dist.barrier()
            ptrs = list(hdl.buffer_ptrs)
            assert len(ptrs) == rt.world_size and all(int(p) != 0 for p in ptrs)
            made.append(PeerExchange(buf, hdl, ptrs, rt.world_size, rt.rank))
        rt.peer, rt.peer_comm = made
        return True
    except Exception as ex:                         # noqa: BLE001 -- any failure means "use NCCL", never a wrong result
        if rt.rank == 0:
            print("scrabble-gan_b200: peer-memory exchange unavailable ({}); small all-reduces use NCCL".format(repr(ex)[:200]),
                  file=sys.stderr)
        rt.peer = rt.peer_comm = None
        return False


def _symm_alloc(rt, numel, dtype):
    """(tensor, rendezvous handle, peer pointers) of a symmetric-memory allocation: a COLLECTIVE call (every rank, same order)."""
    import torch.distributed as dist
    import torch.distributed._symmetric_memory as symm_mem
    group = dist.group.WORLD
    buf = symm_mem.empty(numel, dtype=dtype, device=rt.device)
    hdl = symm_mem.rendezvous(buf, group.group_name)
    ptrs = list(hdl.buffer_ptrs)
    assert len(ptrs) == rt.world_size and all(int(p) != 0 for p in ptrs) and int(ptrs[rt.rank]) == buf.data_ptr()
    return buf, hdl, ptrs


class BucketExchange:
    """A network's flat gradient bucket in symmetric memory + what libsgan's copy-engine all-reduce needs to reduce it
    (sg_peer_bucket_allreduce): the peer pointer table of the bucket and a local staging buffer."""

    def __init__(self, rt, store):
        import ctypes as C
        from . import _abi
        n = store.g.numel()
        g_sym, self.handle, ptrs = _symm_alloc(rt, n, torch.float32)
        g_sym.copy_(store.g)
        store.g = g_sym                                  # Variable.grad views are taken from store.g at access time
        self.ptrs = (C.c_ulonglong * rt.world_size)(*[int(p) for p in ptrs])
        shard = int(_abi.load().sg_peer_bucket_shard(n, rt.world_size))
        self.staging = torch.empty(max(1, (rt.world_size - 1) * shard), device=rt.device, dtype=torch.float32)
        self.n = n
        self.g_ptr = g_sym.data_ptr()


class _StreamWork:
    """Handle of work enqueued on another stream: wait() orders the CURRENT stream after it."""

    def __init__(self, stream):
        self.stream = stream

    def wait(self):
        torch.cuda.current_stream(self.stream.device).wait_stream(self.stream)


def bucket_allreduce_async(rt, store, exposed: bool = False):
    """SUM all-reduce of store.g over the replicas by the copy engines, on the runtime's communication stream, ordered after
    everything enqueued on the current stream; returns a handle whose wait() orders the current stream after the reduction.
    exposed=True: nothing is left to run beside this reduction (the last bucket of a step), so the SMs do the NVLink reads.
    First call per store: moves the bucket into symmetric memory (collective; must happen before any CUDA-graph capture --
    it does: the first steps run eagerly).  None when the peer-memory path is not up (caller falls back to NCCL)."""
    import ctypes as C
    from . import _abi
    if rt.peer is None or rt.peer_comm is None or os.environ.get("SGAN_NO_CE_ALLREDUCE", "0") == "1":
        return None
    ex = getattr(store, "bucket_exchange", None)
    if ex is None or ex.g_ptr != store.g.data_ptr():     # first use (or the bucket was re-allocated): a collective set-up
        if torch.cuda.is_current_stream_capturing():
            return None
        torch.cuda.synchronize(rt.device)
        ex = store.bucket_exchange = BucketExchange(rt, store)
        torch.cuda.synchronize(rt.device)
    comm, ctx = rt.comm_stream_ctx()
    comm.wait_stream(torch.cuda.current_stream(rt.device))
    with torch.cuda.stream(comm):
        _abi.call.sg_peer_bucket_allreduce(ctx, C.c_void_p(store.g.data_ptr()), ex.n, C.c_void_p(ex.staging.data_ptr()), ex.ptrs,
                                           rt.peer_comm.ptrs, rt.world_size, rt.rank, int(exposed and os.environ.get("SGAN_CE_ONLY", "0") != "1"))
    return _StreamWork(comm)


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
