"""Data-parallel equivalence on real GPUs (needs >= 2; run with `gpurun --gpus 2 -- python -m pytest tests/test_dp_gpu.py -m gpu`):
a 2-GPU train step on the two halves of a batch must equal the 1-GPU step on the concatenated batch -- losses /
statistics (global means, global population std for the balancing), BN moving statistics (sync-BN) and the
all-reduced (SUM) gradient buckets of G, D and R.  fp32 mode, tolerance 1e-3 on the whole-gradient rel L2."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _build(rt):
    na = importlib.import_module("scrabble-gan_b200.bigacgan.net_architecture")
    nl = importlib.import_module("scrabble-gan_b200.bigacgan.net_loss")
    optim = importlib.import_module("scrabble-gan_b200.optim")
    G = na.make_generator(128, (32, 160, 1), (32, 8192), None, "B3", 52, vis_model=False, rt=rt, seed=21)
    D = na.make_discriminator((32, 160, 1), None, "B1", vis_model=False, rt=rt, seed=22)
    R = na.make_recognizer((32, 160, 1), None, 53, vis_model=False, rt=rt, seed=23)
    for m in (G, D):                     # exercise attention: sigma != 0
        for v in m.store.vars:
            if v.name.endswith(".sigma"):
                v.assign(np.array([0.1], np.float32))
    gan = na.make_gan(G, D, R, None, vis_model=False)
    opts = optim.setup_optimizer(2e-4, 2e-4, 2e-4, 2e-4, 0.0, 0.999, nl.hinge, 1, 1, 0)
    return G, D, R, gan, opts


def _step(rt, nets, batch):
    du = importlib.import_module("scrabble-gan_b200.bigacgan.data_utils")
    G, D, R, gan, (g_opt, d_opt, r_opt, w_opt, loss_fn, disc_iters, agb) = nets
    imgs, labels, fake, z = batch
    out = du.train_step(0, 0, 1, imgs, labels, D, R, None, gan, g_opt, d_opt, r_opt, w_opt, None, imgs.shape[0], 128, loss_fn,
                        disc_iters, agb, None, 10, "", fake_labels=fake, noise=z)
    grads = {n: m.store.g.detach().cpu().double() for n, m in (("G", G), ("D", D), ("R", R))}
    mov = G.store.s.detach().cpu().double()
    return np.array(out), grads, mov


def _worker(rank, world, port, tmp):
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    runtime = importlib.import_module("scrabble-gan_b200.runtime")
    dp = importlib.import_module("scrabble-gan_b200.dp")
    rt = runtime.Runtime(device=rank, mode="fp32")
    runtime.set_runtime(rt)
    dp.init_data_parallel(rt)
    assert rt.world_size == world
    rng = np.random.RandomState(99)
    B, L = 4, 2
    imgs = rng.uniform(-1, 1, size=(B, 32, 16 * L, 1)).astype(np.float32)
    labels = rng.randint(0, 52, size=(B, L)).astype(np.int32)
    fake = rng.randint(0, 52, size=(B, L)).astype(np.int32)
    z = rng.standard_normal(size=(B, 128)).astype(np.float32)

    nets = _build(rt)
    dp.broadcast_parameters(rt, nets[:3])
    a, b = dp.shard_batch(B, world, rank)
    stats_dp, grads_dp, mov_dp = _step(rt, nets, (imgs[a:b], labels[a:b], fake[a:b], z[a:b]))

    import torch.distributed as dist
    dist.barrier()
    if rank == 0:
        rt.world_size = 1                      # same process, no exchange: the concatenated batch on one GPU
        nets1 = _build(rt)
        stats_1, grads_1, mov_1 = _step(rt, nets1, (imgs, labels, fake, z))
        rt.world_size = world
        for i, (x, y) in enumerate(zip(stats_dp, stats_1)):
            assert abs(x - y) <= 1e-3 * max(abs(y), 0.1), "stat {}: dp {} vs single {}".format(i, x, y)
        # G's gradient runs through the gradient-balancing std over only 4 samples and 7 batch-norms: the fp32 oracle itself
        # sits ~1e-3 from the fp64 one there (tests/test_models_gpu.py), and sharded vs single-pass summation order differs,
        # hence 5e-3 for G and 1e-3 for D and R
        errs = {}
        for n, tol in (("G", 5e-3), ("D", 1e-3), ("R", 1e-3)):
            num = float(((grads_dp[n] - grads_1[n]) ** 2).sum()) ** 0.5
            den = float((grads_1[n] ** 2).sum()) ** 0.5
            errs[n] = (num / max(den, 1e-30), tol, den)
        print("DP-vs-single gradient bucket rel L2:", {k: "%.2e" % v[0] for k, v in errs.items()})
        for n, (e, tol, den) in errs.items():
            assert den > 0 and e <= tol, "{} gradient bucket: rel L2 {} > {}".format(n, e, tol)
        assert float((mov_dp - mov_1).abs().max()) <= 1e-5, "sync-BN moving statistics differ"
        open(os.path.join(tmp, "ok"), "w").write("ok")
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_step_equals_single_gpu_big_batch(tmp_path):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(os.path.join(str(tmp_path), "ok"))


def _peer_worker(rank, world, port, tmp):
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    runtime = importlib.import_module("scrabble-gan_b200.runtime")
    dp = importlib.import_module("scrabble-gan_b200.dp")
    ops = importlib.import_module("scrabble-gan_b200.ops")
    rt = runtime.Runtime(device=rank, mode="fp32")
    runtime.set_runtime(rt)
    dp.init_data_parallel(rt)
    import torch.distributed as dist
    assert rt.peer is not None, "peer-memory exchange did not come up on a 2-GPU NVLink box"
    g = torch.Generator().manual_seed(100 + rank)
    for it, (n, dt) in enumerate([(1, torch.float32), (2048, torch.float32), (16, torch.float64), (4096, torch.float32), (130, torch.float32)] * 3):
        x = torch.randn(n, generator=g, dtype=dt).to(rt.device)
        ref = x.clone()
        dist.all_reduce(ref)
        got = rt.allreduce_small_(x.clone())
        torch.cuda.synchronize()
        assert torch.allclose(got, ref, rtol=1e-6, atol=1e-6), (it, n, dt)
        gathered = [torch.empty_like(got) for _ in range(world)]
        dist.all_gather(gathered, got)
        assert all(torch.equal(gathered[0], t) for t in gathered), "replicas must hold bit-identical sums"
    # gradient buckets: copy-engine reduce-scatter + all-gather over peer memory == NCCL's sum, bit-identical on every replica,
    # repeatable (sequence numbers), also for sizes that do not divide by the world size or by 4, and inside a CUDA graph
    class _Store:
        pass
    for n in (1000003, 5, 4096):
        st = _Store()
        st.g = torch.zeros(n, device=rt.device)
        for rep in range(3):
            vals = torch.randn(n, generator=g).to(rt.device)
            st.g.copy_(vals)
            ref = vals.clone()
            dist.all_reduce(ref)
            h = rt.allreduce_async_(st.g, store=st, exposed=(rep == 1))       # rep 1: the SM pull kernels instead of the copy engines
            assert h is not None and type(h).__name__ == "_StreamWork", "the peer-memory path must be the one that runs"
            h.wait()
            torch.cuda.synchronize()
            assert torch.allclose(st.g, ref, rtol=1e-6, atol=1e-6), (n, rep)
            gathered = [torch.empty_like(st.g) for _ in range(world)]
            dist.all_gather(gathered, st.g)
            assert all(torch.equal(gathered[0], t) for t in gathered), "replicas must hold bit-identical buckets"
    stg = _Store()
    stg.g = torch.zeros(70001, device=rt.device)
    rt.allreduce_async_(stg.g, store=stg).wait()            # first call (moves the bucket to symmetric memory) outside the capture
    torch.cuda.synchronize()
    src = torch.randn(70001, generator=g).to(rt.device)
    graph = torch.cuda.CUDAGraph()
    cap = torch.cuda.Stream(device=rt.device)
    with torch.cuda.graph(graph, stream=cap):
        stg.g.copy_(src)
        rt.allreduce_async_(stg.g, store=stg).wait()
    for rep in range(3):
        graph.replay()
        torch.cuda.synchronize()
        ref = src.clone()
        dist.all_reduce(ref)
        assert torch.allclose(stg.g, ref, rtol=1e-6, atol=1e-6), "bucket all-reduce inside a replayed graph"
    # fused sync-BN statistics == statistics of the concatenated batch
    xs = [torch.randn(3, 8, 10, 64, generator=torch.Generator().manual_seed(7 + r)) * 2 + 0.5 for r in range(world)]
    full = torch.cat(xs, 0).double()
    mm, mv = torch.zeros(64, device=rt.device), torch.ones(64, device=rt.device)
    mean, rstd = ops.bn_stats_finalize_peer(rt, xs[rank].to(rt.device).contiguous(), full.numel() // 64, 64, mm, mv)
    torch.cuda.synchronize()
    m_ref = full.mean((0, 1, 2))
    v_ref = full.var((0, 1, 2), unbiased=False)
    assert torch.allclose(mean.cpu().double(), m_ref, atol=1e-5)
    assert torch.allclose(rstd.cpu().double(), torch.rsqrt(v_ref + 1e-3), rtol=1e-4)
    cnt = full.numel() // 64
    assert torch.allclose(mv.cpu().double(), 0.99 + 0.01 * v_ref * cnt / (cnt - 1), rtol=1e-5)
    dist.barrier()
    if rank == 0:
        open(os.path.join(tmp, "ok_peer"), "w").write("ok")
    dist.destroy_process_group()


def test_peer_memory_small_allreduce_and_fused_sync_bn(tmp_path):
    """csrc/peer.cu: one-shot NVLink peer-memory SUM all-reduce == NCCL, bit-identical on all replicas, and the fused
    stage-2 + exchange + finalize sync-BN launch == statistics of the concatenated batch."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    mp.spawn(_peer_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(os.path.join(str(tmp_path), "ok_peer"))
