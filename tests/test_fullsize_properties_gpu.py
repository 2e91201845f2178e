"""Size-independent properties checked at BASELINE.json's FULL sizes (where the fp64 oracle would take minutes):
integer-valued operands make the bf16 tensor-core convolutions EXACT (products and fp32 partial sums are integers below 2^24),
so linearity must hold bit for bit and spot checks against int64 arithmetic must match exactly; batch independence of D;
softmax-gradient rows of the CTC gradient sum to zero; bit-exact filter-bank indexing at B=64, L=10."""
import importlib

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ops = importlib.import_module("scrabble-gan_b200.ops")
abi = importlib.import_module("scrabble-gan_b200._abi")
na = importlib.import_module("scrabble-gan_b200.bigacgan.net_architecture")
F32, BF16 = abi.SG_F32, abi.SG_BF16


def _ints(rng, shape, lo, hi):
    return torch.from_numpy(rng.randint(lo, hi + 1, size=shape).astype(np.float32))


def test_largest_conv_is_exact_and_linear_on_integer_operands(rt):
    """D.B3.conv2 on the fused [fake;real] batch: N=128, 8x20, 1024 -> 1024, 3x3 (M=20480, K=9216): forward, dgrad, wgrad."""
    rt.set_mode("bf16")
    try:
        rng = np.random.RandomState(0)
        n, h, w, c = 128, 8, 20, 1024
        a, b = _ints(rng, (n, h, w, c), -2, 2), _ints(rng, (n, h, w, c), -2, 2)
        wt = _ints(rng, (3, 3, c, c), -1, 1)
        ad, bd, sd = (t.to(rt.device).to(torch.bfloat16) for t in (a, b, a + b))
        wd = wt.to(rt.device)
        d = ops.desc_conv_fwd(n, h, w, c, c, 3, 3, "same", BF16, F32)
        wp = ops.pack_weights(rt, d, wd)
        outs = []
        for x in (ad, bd, sd):
            o = rt.empty((n, h, w, c))
            ops.conv_run(rt, d, x, wd, wp, None, None, o)
            outs.append(o)
        assert torch.equal(outs[0] + outs[1], outs[2]), "conv(a) + conv(b) != conv(a + b) on exactly representable operands"
        # spot checks against int64 arithmetic (SAME padding: taps outside the image contribute 0)
        A, W = a.to(torch.int64), wt.to(torch.int64)
        got = outs[0].cpu()
        for (ni, y, x_, co) in ((0, 0, 0, 0), (5, 3, 7, 100), (127, 7, 19, 1023), (64, 4, 0, 511), (17, 0, 19, 3)):
            acc = 0
            for dy in (-1, 0, 1):
                for dx in (-1, 0, 1):
                    yy, xx = y + dy, x_ + dx
                    if 0 <= yy < h and 0 <= xx < w:
                        acc += int((A[ni, yy, xx] * W[dy + 1, dx + 1, :, co]).sum())
            assert int(got[ni, y, x_, co]) == acc
        # dgrad is the same kernel with the channel roles swapped: linear and exact as well
        dd = ops.desc_conv_dgrad(n, h, w, c, c, 3, 3, "same", BF16, F32)
        wpd = ops.pack_weights(rt, dd, wd)
        gs = []
        for x in (ad, bd, sd):
            o = rt.empty((n, h, w, c))
            ops.conv_run(rt, dd, x, wd, wpd, None, None, o)
            gs.append(o)
        assert torch.equal(gs[0] + gs[1], gs[2])
        # filter gradient: dW[a,b,ci,co] = sum_pixels x[p + tap, ci] * dy[p, co]: exact integers (|sum| < 2^24)
        dy = _ints(rng, (n, h, w, c), -1, 1)
        dyd = dy.to(rt.device).to(torch.bfloat16)
        dw = rt.zeros((3, 3, c, c))
        dwg = ops.desc_conv_fwd(n, h, w, c, c, 3, 3, "same", BF16, BF16)
        ops.conv_wgrad(rt, dwg, ad, dyd, dw)
        dwc = dw.cpu()
        DY = dy.to(torch.int64)
        for (ta, tb, ci, co) in ((1, 1, 0, 0), (0, 0, 17, 900), (2, 2, 1023, 1023), (0, 2, 512, 1), (2, 1, 5, 640)):
            dy_, dx_ = ta - 1, tb - 1
            ys = slice(max(0, -dy_), h - max(0, dy_))          # output rows whose tap stays inside the image
            xs = slice(max(0, -dx_), w - max(0, dx_))
            xin = A[:, ys.start + dy_:ys.stop + dy_, xs.start + dx_:xs.stop + dx_, ci]
            exp = int((xin * DY[:, ys, xs, co]).sum())
            assert int(dwc[ta, tb, ci, co]) == exp
    finally:
        rt.set_mode("fp32")


def test_discriminator_is_batch_independent_at_full_size(rt):
    """D has no cross-sample coupling: logits of a 64-image batch == logits of its two halves (what the fused [fake;real]
    pass and the data-parallel sharding rely on)."""
    rt.set_mode("bf16")
    try:
        D = na.make_discriminator((32, 160, 1), None, "B1", vis_model=False, rt=rt, seed=7)
        for v in D.store.vars:
            if v.name.endswith(".sigma"):
                v.assign(np.array([0.1], np.float32))
        x = torch.from_numpy(np.random.RandomState(1).uniform(-1, 1, size=(64, 32, 80, 1)).astype(np.float32)).to(rt.device)
        full, _ = D.forward(rt, x)
        lo, _ = D.forward(rt, x[:32].contiguous())
        hi, _ = D.forward(rt, x[32:].contiguous())
        halves = torch.cat([lo.view(-1), hi.view(-1)])
        scale = float(full.abs().max())
        assert float((full.view(-1) - halves).abs().max()) <= 1e-5 * max(scale, 1.0)
    finally:
        rt.set_mode("fp32")


def test_ctc_gradient_rows_sum_to_zero_at_config3_size(rt):
    """BASELINE configs[2]: B=256, T=39, C=81, L=10.  d loss / d logits = (softmax - posterior): every frame's row sums to 0,
    the loss is positive and finite, and shuffling the batch permutes the per-sample losses."""
    rng = np.random.RandomState(2)
    b, t, c, l = 256, 39, 81, 10
    logits = torch.from_numpy((rng.standard_normal((b, t, c)) * 2).astype(np.float32)).to(rt.device)
    labels = torch.from_numpy(rng.randint(0, c - 1, size=(b, l)).astype(np.int32)).to(rt.device)
    loss, grad = ops.ctc(rt, logits, labels)
    assert torch.isfinite(loss).all() and float(loss.min()) > 0
    assert float(grad.sum(-1).abs().max()) <= 1e-4
    perm = torch.from_numpy(rng.permutation(b)).to(rt.device)
    loss_p, _ = ops.ctc(rt, logits[perm].contiguous(), labels[perm].contiguous())
    assert torch.allclose(loss_p, loss[perm], rtol=1e-6, atol=1e-6)


def test_filter_bank_indexing_is_bit_exact_at_full_size(rt):
    """B=64, L=10, vocab 52: with a one-hot z the output must be single bank rows, unrounded, at
    out[b, k%4, 4l + k//2048, (k%2048)//4] = bank[y[b,l], j, k]."""
    rng = np.random.RandomState(3)
    b, l, vocab, j = 64, 10, 52, 13
    bank = torch.from_numpy(rng.standard_normal((vocab, 32, 8192)).astype(np.float32))
    y = torch.from_numpy(rng.randint(0, vocab, size=(b, l)).astype(np.int32))
    z = torch.zeros(b, 32)
    z[:, j] = 1.0
    out = ops.filterbank_fwd(rt, z.to(rt.device), 32, y.to(rt.device), bank.to(rt.device)).cpu()
    rows = bank[y.long(), j]                                   # (b, l, 8192)
    k = torch.arange(8192)
    exp = torch.zeros(b, 4, 4 * l, 512)
    for li in range(l):
        exp[:, k % 4, 4 * li + k // 2048, (k % 2048) // 4] = rows[:, li]
    assert torch.equal(out, exp)


def test_run_inference_path_accepts_a_15_character_word(rt):
    """run_inference.py:27-35 encodes 'machinelearning' (15 characters > the training buckets) and calls the generator with
    [noise, labels], training=False: the fully-convolutional G must take any length (32 x 240 image; the 7680 x 1920
    attention map exceeds the tensor-op kernel's shared memory and takes the exact FFMA path in bf16 mode too)."""
    char_vec = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ"
    word = "machinelearning"
    for mode in ("bf16", "fp32"):
        rt.set_mode(mode)
        try:
            G = na.make_generator(128, (32, 160, 1), (32, 8192), None, "B3", 52, vis_model=False, rt=rt, seed=5)
            for v in G.store.vars:
                if v.name.endswith(".sigma"):
                    v.assign(np.array([0.2], np.float32))
            labels = np.array([[char_vec.index(ch) for ch in word]] * 3, np.int32)
            noise = np.random.RandomState(0).standard_normal((3, 128)).astype(np.float32)
            img = G([noise, labels], training=False)
            assert tuple(img.shape) == (3, 32, 16 * len(word), 1)
            assert torch.isfinite(img).all() and float(img.min()) >= -1.0 and float(img.max()) <= 1.0
            again = G([noise, labels], training=False)
            assert torch.equal(img, again), "inference must be deterministic"
            pred = ((img + 1) / 2.0).cpu().numpy()                 # run_inference.py:37
            assert pred.min() >= 0.0 and pred.max() <= 1.0
        finally:
            rt.set_mode("fp32")


def test_tf32_conv_is_exact_on_integer_operands(rt):
    """Same exactness argument for the kind::tf32 path (fp32 storage): D.B2.conv2 shape, N=16, 16x40, 512 -> 512."""
    rt.set_mode("tf32")
    try:
        rng = np.random.RandomState(4)
        n, h, w, c = 16, 16, 40, 512
        a, b = _ints(rng, (n, h, w, c), -3, 3), _ints(rng, (n, h, w, c), -3, 3)
        wt = _ints(rng, (3, 3, c, c), -1, 1)
        d = ops.desc_conv_fwd(n, h, w, c, c, 3, 3, "same", F32, F32)
        assert ops.tc_ok(rt, d)
        wd = wt.to(rt.device)
        wp = ops.pack_weights(rt, d, wd)
        outs = []
        for x in (a, b, a + b):
            o = rt.empty((n, h, w, c))
            ops.conv_run(rt, d, x.to(rt.device), wd, wp, None, None, o)
            outs.append(o)
        assert torch.equal(outs[0] + outs[1], outs[2])
        A, W = a.to(torch.int64), wt.to(torch.int64)
        got = outs[0].cpu()
        for (ni, y, x_, co) in ((0, 0, 0, 0), (15, 15, 39, 511), (7, 8, 20, 300)):
            acc = 0
            for dy in (-1, 0, 1):
                for dx in (-1, 0, 1):
                    yy, xx = y + dy, x_ + dx
                    if 0 <= yy < h and 0 <= xx < w:
                        acc += int((A[ni, yy, xx] * W[dy + 1, dx + 1, :, co]).sum())
            assert int(got[ni, y, x_, co]) == acc
    finally:
        rt.set_mode("fp32")


def test_fork_mode_step_at_the_gin_batch_size(rt):
    """Mode B ("as-written fork": in-G style encoder on (32,160,1) style images + style promoter W) at the reference's own
    configuration (scrabble_gan.gin: batch 16, hinge, words <= 10 characters): three steps run, every statistic is finite,
    the optimizers of all four networks advance, and weights change."""
    du = importlib.import_module("scrabble-gan_b200.bigacgan.data_utils")
    nl = importlib.import_module("scrabble-gan_b200.bigacgan.net_loss")
    optim = importlib.import_module("scrabble-gan_b200.optim")
    rt.set_mode("bf16")
    try:
        in_dim = (32, 160, 1)
        G = na.make_generator(128, in_dim, (32, 8192), None, "B3", 52, vis_model=False, style_encoder=True, rt=rt, seed=51)
        D = na.make_discriminator(in_dim, None, "B1", vis_model=False, rt=rt, seed=52)
        R = na.make_recognizer(in_dim, None, 53, vis_model=False, rt=rt, seed=53)
        W = na.make_style_promoter(in_dim, None, "B1", vis_model=False, rt=rt, seed=54)
        gan = na.make_gan(G, D, R, W, vis_model=False)
        g_opt, d_opt, r_opt, w_opt, loss_fn, disc_iters, agb = optim.setup_optimizer(2e-4, 2e-4, 2e-4, 2e-4, 0.0, 0.999, nl.hinge, 1, 0, 0)
        rng = np.random.RandomState(6)
        bs = 16
        words = du.synthetic_random_words(10, "x" * 52, words_per_bucket=32, seed=1)
        w0 = {n: m.store.w.clone() for n, m in (("G", G), ("D", D), ("R", R), ("W", W))}
        for i, length in enumerate((5, 10, 3)):
            imgs = rng.uniform(-1, 1, size=(bs, 32, 16 * length, 1)).astype(np.float32)
            labels = rng.randint(0, 52, size=(bs, length)).astype(np.int32)
            style = [rng.uniform(-1, 1, size=(32, 160)).astype(np.float32) for _ in range(bs)]       # list of B (32,160) images
            out = du.train_step(0, i, 3, imgs, labels, D, R, W, gan, g_opt, d_opt, r_opt, w_opt, style, bs, 128, loss_fn, disc_iters,
                                agb, words, 10, "")
            assert len(out) == 16 and all(np.isfinite(v) for v in out), out
        assert (g_opt.iterations, d_opt.iterations, r_opt.iterations, w_opt.iterations) == (3, 3, 3, 3)
        for n, m in (("G", G), ("D", D), ("R", R), ("W", W)):
            assert not torch.equal(m.store.w, w0[n]), n + " did not move"
    finally:
        rt.set_mode("fp32")
