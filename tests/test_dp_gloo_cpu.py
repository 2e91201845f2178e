"""Host-side data-parallel logic on CPU with torch.distributed `gloo`, world_size 2 (no GPU needed).

Covers what the N>1 path does around the kernels (scrabble-gan_b200/dp.py, Runtime.allreduce_): rendezvous from the
torchrun environment, rank-0 parameter broadcast, SUM (not mean) gradient all-reduce over the flat bucket
(SURVEY Q7), even batch sharding, a length schedule shared by all ranks, and -- with the CPU oracle standing in for
the kernels -- that N replicas on shards reduce to one replica on the concatenated batch for the three exchange
points of SURVEY.md section 8e: parameter gradients, batch-norm raw sums, gradient-balance sums."""
import importlib
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _StubStore:
    def __init__(self, n, fill):
        self.w = torch.full((n,), float(fill))
        self.s = torch.full((4,), float(fill) + 0.5)
        self.g = torch.zeros(n)
        self.version = 0


class _StubModel:
    def __init__(self, n, fill):
        self.store = _StubStore(n, fill)


class _StubRuntime:
    """The CPU-visible part of runtime.Runtime used by dp.py (the real one refuses to exist without CUDA)."""
    def __init__(self):
        self.device = torch.device("cpu")
        self.world_size, self.rank, self.process_group = 1, 0, None


def _worker(rank, world, port, tmp):
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(2)
    dp = importlib.import_module("scrabble-gan_b200.dp")
    runtime = importlib.import_module("scrabble-gan_b200.runtime")
    import sgan_oracle as O
    rt = _StubRuntime()
    dp.init_data_parallel(rt, backend="gloo")
    assert (rt.world_size, rt.rank) == (world, rank)
    allreduce = lambda t: runtime.Runtime.allreduce_(rt, t)

    # (0) broadcast: every replica starts from rank 0's weights and BN moving statistics
    m = _StubModel(10, fill=rank + 1)
    dp.broadcast_parameters(rt, [m])
    assert float(m.store.w[0]) == 1.0 and float(m.store.s[0]) == 1.5 and m.store.version == 1

    # shared length schedule and even shards
    assert dp.length_schedule(17) == dp.length_schedule(17) and 1 <= dp.length_schedule(17)[0] <= 10
    sched = torch.tensor(dp.length_schedule(5) + dp.length_schedule(6))
    other = sched.clone()
    dist.broadcast(other, src=0)
    assert torch.equal(sched, other)
    a, b = dp.shard_batch(8, world, rank)
    assert (a, b) == (rank * 4, rank * 4 + 4)
    with pytest.raises(AssertionError):
        dp.shard_batch(7, world, rank)

    # (3) parameter gradients: SUM all-reduce of per-shard gradients of per-shard loss SUMS == big-batch gradient
    g = torch.Generator().manual_seed(5)
    B = 4
    x = torch.rand(B, 32, 16, 1, generator=g, dtype=torch.float64) * 2 - 1
    w = {k: v for k, v in O.make_recognizer_params(3, torch.float64).items()}
    labels = torch.randint(0, 52, (B, 1), generator=g)

    def r_grad(xs, ys):
        leaf = {k: v.clone().requires_grad_(not k.endswith(O.NON_TRAINABLE_SUFFIXES)) for k, v in w.items()}
        n = xs.shape[0]
        loss = O.recognizer(xs, ys, torch.full((n, 1), 3), torch.full((n, 1), 1), leaf)
        names = O.trainable_names(leaf)
        gs = torch.autograd.grad(loss.sum(), [leaf[k] for k in names])
        return torch.cat([t.reshape(-1) for t in gs]), loss.detach()
    full, loss_full = r_grad(x, labels)
    mine, loss_mine = r_grad(x[a // 2:b // 2], labels[a // 2:b // 2])     # B=4 over 2 ranks: 2 images each
    bucket = mine.clone()
    allreduce(bucket)
    assert float((bucket - full).abs().max()) <= 1e-9 * float(full.abs().max())

    # (1) batch-norm raw sums (sum x, sum x^2) all-reduced, count * world == statistics of the concatenated batch
    act = torch.randn(B, 4, 6, 8, generator=g, dtype=torch.float64) * 3 + 1
    shard = act[a // 2:b // 2]
    sums = torch.cat([shard.sum((0, 1, 2)), (shard * shard).sum((0, 1, 2))])
    allreduce(sums)
    count = shard.numel() // 8 * world
    mean = sums[:8] / count
    var = sums[8:] / count - mean * mean
    xh, nm, nv = O.batchnorm_train(act, torch.zeros(8, dtype=torch.float64), torch.ones(8, dtype=torch.float64))
    assert torch.allclose((shard - mean) * torch.rsqrt(var + O.BN_EPS), xh[a // 2:b // 2], atol=1e-10)
    assert torch.allclose(nv, 0.99 + 0.01 * var * count / (count - 1), atol=1e-12)

    # (2) gradient-balance sums (sum g, sum g^2, sum r, sum r^2, N) -> population std over the GLOBAL batch
    r = torch.rand(B, 1, generator=g, dtype=torch.float64) * 10
    gl = torch.randn(B, 1, generator=g, dtype=torch.float64)
    rs, gs_ = r[a // 2:b // 2], gl[a // 2:b // 2]
    five = torch.tensor([gs_.sum(), (gs_ * gs_).sum(), rs.sum(), (rs * rs).sum(), float(rs.numel())], dtype=torch.float64)
    allreduce(five)
    n = five[4]
    g_std = (five[1] / n - (five[0] / n) ** 2).sqrt()
    r_std = (five[3] / n - (five[2] / n) ** 2).sqrt()
    gb, rb, _, r_std_ref, g_std_ref = O.apply_gradient_balancing(r, gl, 1.0)
    assert abs(float(g_std - g_std_ref)) <= 1e-12 and abs(float(r_std - r_std_ref)) <= 1e-12
    assert torch.allclose(gs_ + (g_std / r_std) * rs, gb[a // 2:b // 2], atol=1e-12)

    dist.barrier()
    dist.destroy_process_group()
    open(os.path.join(tmp, "ok%d" % rank), "w").write("ok")


def test_world_size_2_gloo(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(os.path.join(str(tmp_path), "ok%d" % r)) for r in range(world))
