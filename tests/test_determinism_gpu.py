"""Bitwise repeatability of the gradients (deterministic split-K): the filter-gradient kernels add their pixel
splits in split order (turn semaphores) and every other cross-block sum goes through per-block partials that the last
block adds in block order (csrc/common.cuh), so two runs from the same state must agree BIT FOR BIT -- which is also what
makes Adam's first steps (lr * sign(g) with beta1 = 0) reproducible."""
import importlib

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ops = importlib.import_module("scrabble-gan_b200.ops")
abi = importlib.import_module("scrabble-gan_b200._abi")
na = importlib.import_module("scrabble-gan_b200.bigacgan.net_architecture")
du = importlib.import_module("scrabble-gan_b200.bigacgan.data_utils")
nl = importlib.import_module("scrabble-gan_b200.bigacgan.net_loss")
optim = importlib.import_module("scrabble-gan_b200.optim")
F32, BF16 = abi.SG_F32, abi.SG_BF16
IN_DIM = (32, 160, 1)


@pytest.mark.parametrize("shape", [(128, 16, 40, 512, 512, 3), (128, 16, 40, 64, 512, 3), (64, 8, 20, 512, 1024, 3), (32, 32, 80, 64, 64, 3),
                                   (16, 16, 40, 64, 512, 1)])
def test_wgrad_tc_is_bitwise_repeatable(rt, shape):
    """Layer shapes of D at the benchmark size whose filter gradient is split over pixel ranges (splits > 1): five runs, one
    answer; and the answer does not depend on what the gradient buffer held before (it is accumulated into)."""
    rt.set_mode("bf16")
    try:
        n, h, w, ci, co, k = shape
        g = torch.Generator(device=rt.device).manual_seed(3)
        x = torch.randn(n, h, w, ci, generator=g, device=rt.device).bfloat16()
        dy = torch.randn(n, h, w, co, generator=g, device=rt.device).bfloat16()
        d = ops.desc_conv_fwd(n, h, w, ci, co, k, k, "same", BF16, BF16)
        outs = []
        for _ in range(5):
            dw = torch.zeros(k, k, ci, co, device=rt.device)
            ops.conv_wgrad(rt, d, x, dy, dw)
            outs.append(dw)
        rt.sync()
        for o in outs[1:]:
            assert torch.equal(outs[0], o), "filter gradient differs between runs (max diff {:.3e})".format(float((outs[0] - o).abs().max()))
        ref = torch.einsum("nhwi,nhwo->io", x[:, :, :, :].float(), dy.float()) if k == 1 else None
        if ref is not None:
            assert float((outs[0][0, 0] - ref).abs().max()) <= 2e-3 * float(ref.abs().max())
    finally:
        rt.set_mode("fp32")


@pytest.mark.parametrize("rows,cols,dt", [(128 * 32 * 80, 64, BF16), (128 * 16 * 40, 512, F32), (4099, 24, F32), (64 * 39, 53, F32)])
def test_cross_block_sums_are_bitwise_repeatable(rt, rows, cols, dt):
    g = torch.Generator(device=rt.device).manual_seed(4)
    x = torch.randn(rows, cols, generator=g, device=rt.device)
    x = x.bfloat16() if dt == BF16 else x
    outs = []
    for _ in range(4):
        out = torch.full((cols,), 0.5, device=rt.device)
        ops.colsum_into(rt, x, cols, out, accumulate=1)
        outs.append(out)
    rt.sync()
    for o in outs[1:]:
        assert torch.equal(outs[0], o)
    exp = x.double().sum(0) + 0.5
    assert float((outs[0].double() - exp).abs().max()) <= 1e-4 * float(exp.abs().max() + 1)
    a, b = torch.randn(1 << 20, generator=g, device=rt.device), torch.randn(1 << 20, generator=g, device=rt.device)
    dots = []
    for _ in range(4):
        o = torch.full((1,), 2.0, device=rt.device)
        ops.dot_into(rt, a, b, o, accumulate=1)
        dots.append(o)
    rt.sync()
    assert all(torch.equal(dots[0], o) for o in dots[1:])
    assert abs(float(dots[0]) - 2.0 - float((a.double() * b.double()).sum())) <= 1e-3 * (1 << 10)


def _fresh(rt, seed):
    G = na.make_generator(128, IN_DIM, (32, 8192), None, "B3", 52, vis_model=False, rt=rt, seed=seed + 1)
    D = na.make_discriminator(IN_DIM, None, "B1", vis_model=False, rt=rt, seed=seed + 2)
    R = na.make_recognizer(IN_DIM, None, 53, vis_model=False, rt=rt, seed=seed + 3)
    for m in (G, D):
        for v in m.store.vars:
            if v.name.endswith(".sigma"):
                v.assign(np.array([0.1], np.float32))
    return G, D, R


@pytest.mark.parametrize("mode,b,l_r,l_f", [("bf16", 16, 5, 5), ("bf16", 8, 3, 4), ("tf32", 8, 4, 4), ("fp32", 4, 2, 2)])
def test_train_steps_are_bitwise_repeatable(rt, mode, b, l_r, l_f):
    """Three train steps, twice, from identically initialised models: every gradient bucket after the last step and every
    weight must be identical bit for bit (fused and unfused D/R batches; eager)."""
    rt.set_mode(mode)
    old = du.GRAPH_ENABLED
    try:
        du._graph_cache.clear()
        du.GRAPH_ENABLED = False
        rng = np.random.RandomState(11)
        batches = [(rng.uniform(-1, 1, size=(b, 32, 16 * l_r, 1)).astype(np.float32), rng.randint(0, 52, size=(b, l_r)).astype(np.int32),
                    rng.randint(0, 52, size=(b, l_f)).astype(np.int32), rng.standard_normal(size=(b, 128)).astype(np.float32)) for _ in range(3)]
        results = []
        for _ in range(2):
            G, D, R = _fresh(rt, 70)
            gan = na.make_gan(G, D, R, None, vis_model=False)
            g_opt, d_opt, r_opt, w_opt, loss_fn, disc_iters, agb = optim.setup_optimizer(2e-4, 2e-4, 2e-4, 2e-4, 0.0, 0.999, nl.hinge, 1, 1, 0)
            stats = []
            for i, (imgs, labels, fake, z) in enumerate(batches):
                stats.append(du.train_step(0, i, 3, imgs, labels, D, R, None, gan, g_opt, d_opt, r_opt, w_opt, None, b, 128, loss_fn, disc_iters,
                                           agb, None, 10, "", fake_labels=fake, noise=z))
            results.append((stats, [m.store.g.clone() for m in (G, D, R)], [m.store.w.clone() for m in (G, D, R)],
                            [m.store.s.clone() for m in (G, D, R)]))
        (s0, g0, w0, m0), (s1, g1, w1, m1) = results
        assert s0 == s1, "the 16 statistics differ between two identical runs"
        for name, a, c in zip("GDR", g0, g1):
            assert torch.equal(a, c), "{} gradients differ between runs: {} of {} entries".format(name, int((a != c).sum()), a.numel())
        for name, a, c in zip("GDR", w0, w1):
            assert torch.equal(a, c), "{} weights differ between runs: {} of {} entries".format(name, int((a != c).sum()), a.numel())
        for a, c in zip(m0, m1):
            assert torch.equal(a, c)
    finally:
        du.GRAPH_ENABLED = old
        du._graph_cache.clear()
        rt.set_mode("fp32")
