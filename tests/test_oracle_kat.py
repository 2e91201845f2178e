"""Math-only known-answer tests that pin the CPU oracle WITHOUT depending on the restatement itself (the reference
ships no tests or golden vectors -- SURVEY.md section 4): brute-force CTC path enumeration, 7-loop convolution,
explicit scatter for the transposed convolutions, SVD for spectral norm, finite differences for the closed-form
gradient-balance derivative, exhaustive filter-bank index map."""
import math

import numpy as np
import torch

import sgan_oracle as O

DT = torch.float64


def test_ctc_matches_brute_force_enumeration():
    g = torch.Generator().manual_seed(0)
    for t_len, c, labels in ((4, 3, [0]), (5, 4, [1, 1]), (5, 4, [0, 2]), (6, 3, [1, 0, 1]), (3, 5, [2, 3])):
        probs = torch.softmax(torch.randn(1, t_len, c, generator=g, dtype=DT), -1)
        got = O.ctc_batch_cost(torch.tensor([labels]), probs, torch.tensor([[t_len]]), torch.tensor([[len(labels)]]))
        q = (probs[0] + 1e-7) / (1 + c * 1e-7)          # Keras adds eps, tf.nn.ctc_loss re-softmaxes
        exp = O.ctc_brute_force(q, labels, c - 1)
        assert abs(float(got) - exp) < 1e-10 * max(1.0, abs(exp))


def test_conv2d_same_and_valid_match_direct_loops():
    g = torch.Generator().manual_seed(1)
    for (h, w, k, pad) in ((5, 7, 3, "same"), (4, 6, 1, "same"), (2, 9, 2, "valid"), (6, 5, 2, "same")):
        x = torch.randn(2, h, w, 3, generator=g, dtype=DT)
        wt = torch.randn(k, k, 3, 4, generator=g, dtype=DT)
        b = torch.randn(4, generator=g, dtype=DT)
        got = O.conv2d(x, wt, b, pad)
        if pad == "same":
            pt = pl = (k - 1) // 2
            oh, ow = h, w
        else:
            pt = pl = 0
            oh, ow = h - k + 1, w - k + 1
        exp = O.conv2d_loops(x.numpy(), wt.numpy(), b.numpy(), pt, pl, oh, ow)
        assert np.abs(got.numpy() - exp).max() < 1e-12


def test_conv2d_transpose_follows_tf_same_rule():
    """y = 2i + kh (pad_before 0, output cropped to 2n) for stride 2; x = j + kw - 1 for stride 1; 1x1 stride 2 writes
    only even positions; bias everywhere."""
    g = torch.Generator().manual_seed(2)
    for (k, sy, sx) in ((3, 2, 2), (3, 2, 1), (1, 2, 2), (1, 2, 1)):
        n, h, w, ci, co = 2, 3, 4, 2, 3
        x = torch.randn(n, h, w, ci, generator=g, dtype=DT)
        wt = torch.randn(k, k, co, ci, generator=g, dtype=DT)
        b = torch.randn(co, generator=g, dtype=DT)
        got = O.conv2d_transpose(x, wt, b, (sy, sx)).numpy()
        exp = np.zeros((n, h * sy, w * sx, co))
        pby, pbx = max(k - sy, 0) // 2, max(k - sx, 0) // 2
        for ni in range(n):
            for i in range(h):
                for j in range(w):
                    for a in range(k):
                        for c in range(k):
                            yy, xx = i * sy + a - pby, j * sx + c - pbx
                            if 0 <= yy < h * sy and 0 <= xx < w * sx:
                                exp[ni, yy, xx] += wt[a, c].numpy() @ x[ni, i, j].numpy()
        exp += b.numpy()
        assert np.abs(got - exp).max() < 1e-12


def test_filter_bank_index_map_is_exhaustive():
    g = torch.Generator().manual_seed(3)
    b, l, vocab = 2, 3, 5
    bank = torch.randn(vocab, 32, 8192, generator=g, dtype=DT)
    y = torch.randint(0, vocab, (b, l), generator=g)
    z0 = torch.randn(b, 32, generator=g, dtype=DT)
    out = O.filter_bank(z0, y, bank)
    assert out.shape == (b, 4, 4 * l, 512)
    raw = torch.einsum("bj,bljk->blk", z0, bank[y])
    k = torch.arange(8192)
    for li in range(l):
        hh, ww, cc = k % 4, 4 * li + k // 2048, (k % 2048) // 4
        assert torch.equal(out[:, hh, ww, cc], raw[:, li, :])
        assert O.filter_bank_index_map(li, 4099) == (4099 % 4, 4 * li + 2, (4099 % 2048) // 4)
    # the map is a bijection onto the (4, 4L, 512) grid
    seen = {O.filter_bank_index_map(li, kk) for li in range(l) for kk in range(8192)}
    assert len(seen) == l * 8192


def test_batchnorm_train_normalises_and_updates_moving_stats():
    g = torch.Generator().manual_seed(4)
    x = torch.randn(4, 3, 5, 8, generator=g, dtype=DT) * 2 + 1
    xh, mm, mv = O.batchnorm_train(x, torch.zeros(8, dtype=DT), torch.ones(8, dtype=DT))
    assert xh.mean((0, 1, 2)).abs().max() < 1e-12
    var = x.var((0, 1, 2), unbiased=False)
    assert ((xh.var((0, 1, 2), unbiased=False) - var / (var + 1e-3)).abs().max()) < 1e-12
    assert torch.allclose(mm, 0.01 * x.mean((0, 1, 2)))
    assert torch.allclose(mv, 0.99 + 0.01 * x.var((0, 1, 2), unbiased=True))


def test_spectral_norm_converges_to_largest_singular_value():
    g = torch.Generator().manual_seed(5)
    w = torch.randn(3, 3, 4, 6, generator=g, dtype=DT)
    u = torch.randn(1, 6, generator=g, dtype=DT)
    wn = O.spectral_norm(w, u, power_iteration=200)
    s = torch.linalg.svdvals(w.reshape(-1, 6))
    assert abs(float(torch.linalg.svdvals(wn.reshape(-1, 6))[0]) - 1.0) < 1e-8
    assert torch.allclose(wn * s[0], w, atol=1e-8)
    # one iteration (the reference's setting) is an under-estimate of sigma_max
    w1 = O.spectral_norm(w, u, 1)
    assert float(torch.linalg.svdvals(w1.reshape(-1, 6))[0]) >= 1.0 - 1e-12


def test_gradient_balance_closed_form_derivative():
    """dS/dg_i = 1 + (R/sd_r)(g_i - mean_g)/(N sd_g);  dS/dr_i = sd_g/sd_r - sd_g R (r_i - mean_r)/(N sd_r^3)."""
    g = torch.Generator().manual_seed(6)
    n = 7
    r = (torch.randn(n, 1, generator=g, dtype=DT) + 20).requires_grad_(True)
    gl = torch.randn(n, 1, generator=g, dtype=DT).requires_grad_(True)
    gb, rb, alpha, sr, sg = O.apply_gradient_balancing(r, gl, 1.0)
    gb.sum().backward()
    R, N = float(r.sum()), n
    dg = 1 + (R / sr) * (gl - gl.mean()) / (N * sg)
    dr = sg / sr - sg * R * (r - r.mean()) / (N * sr ** 3)
    assert torch.allclose(gl.grad, dg.detach(), atol=1e-10)
    assert torch.allclose(r.grad, dr.detach(), atol=1e-10)
    # and against central finite differences
    eps = 1e-6
    for i in (0, 3):
        rp, rm = r.detach().clone(), r.detach().clone()
        rp[i] += eps
        rm[i] -= eps
        fd = (O.apply_gradient_balancing(rp, gl.detach())[0].sum() - O.apply_gradient_balancing(rm, gl.detach())[0].sum()) / (2 * eps)
        assert abs(float(fd) - float(r.grad[i])) < 1e-5


def test_losses_and_adam_formulas():
    d_real, d_fake = torch.tensor([[0.3], [2.0]], dtype=DT), torch.tensor([[-0.5], [-3.0]], dtype=DT)
    z = torch.zeros(2, 1, dtype=DT)
    d_loss, dlr, dlf, g_loss, *_ = O.hinge(d_real, d_fake, z, z)
    assert torch.allclose(dlr, torch.tensor([[0.7], [0.0]], dtype=DT)) and torch.allclose(dlf, torch.tensor([[0.5], [0.0]], dtype=DT))
    assert torch.allclose(g_loss, -d_fake)
    x = torch.tensor([0.7, -1.3], dtype=DT)
    assert torch.allclose(O._sce(x, True), -torch.log(torch.sigmoid(x))) and torch.allclose(O._sce(x, False), -torch.log(1 - torch.sigmoid(x)))
    w, g, m, v = (torch.tensor([x], dtype=DT) for x in (1.0, 0.5, 0.0, 0.0))
    w1, m1, v1 = O.adam_update(w, g, m, v, 1, lr=2e-4, beta1=0.0, beta2=0.999)
    lr_t = 2e-4 * math.sqrt(1 - 0.999)
    assert abs(float(w1) - (1.0 - lr_t * 0.5 / (math.sqrt(0.001 * 0.25) + 1e-7))) < 1e-9


def test_parameter_counts_match_the_reference_models():
    """D = W = 37 336 384 (+1 sigma, +5 120 attention projections kept persistent here);
    filter bank 13 631 488; G-core 2 582 530; R 5 578 037 (53 outputs)."""
    def count(p):
        return sum(v.numel() for k, v in p.items() if not k.endswith(O.NON_TRAINABLE_SUFFIXES))
    attn = 64 * 8 * 2 + 64 * 32 + 32 * 64
    assert count(O.make_discriminator_params(0, torch.float32)) == 37336384 + 1 + attn
    assert count(O.make_recognizer_params(0, torch.float32)) == 5578037
    gp = O.make_generator_params(0, torch.float32)
    assert gp["filter_bank"].numel() == 13631488
    assert count(gp) - gp["filter_bank"].numel() == 2582530 + attn


def test_non_local_block_matches_per_pixel_loops():
    """NonLocalBlock (arch_ops.py:32-67) written out with explicit loops in numpy: 1x1 projections, 2x2 max-pool of phi / g,
    softmax over the pooled keys WITHOUT 1/sqrt(d), projection back, sigma * o + x."""
    g = torch.Generator().manual_seed(31)
    n, h, w, c = 2, 4, 6, 16
    x = torch.randn(n, h, w, c, generator=g, dtype=DT)
    p = {"a.theta.w": torch.randn(1, 1, c, c // 8, generator=g, dtype=DT), "a.phi.w": torch.randn(1, 1, c, c // 8, generator=g, dtype=DT),
         "a.g.w": torch.randn(1, 1, c, c // 2, generator=g, dtype=DT), "a.o.w": torch.randn(1, 1, c // 2, c, generator=g, dtype=DT),
         "a.sigma": torch.tensor([0.3], dtype=DT)}
    got = O.non_local_block(x, p, "a").numpy()
    X = x.numpy()
    Wt, Wp, Wg, Wo = (p[k][0, 0].numpy() for k in ("a.theta.w", "a.phi.w", "a.g.w", "a.o.w"))
    exp = np.zeros_like(X)
    for ni in range(n):
        theta = X[ni] @ Wt                                     # (h, w, c/8)
        phi_f, g_f = X[ni] @ Wp, X[ni] @ Wg
        phi = phi_f.reshape(h // 2, 2, w // 2, 2, -1).max(axis=(1, 3)).reshape(-1, c // 8)
        gg = g_f.reshape(h // 2, 2, w // 2, 2, -1).max(axis=(1, 3)).reshape(-1, c // 2)
        for i in range(h):
            for j in range(w):
                s = phi @ theta[i, j]
                a = np.exp(s - s.max())
                a /= a.sum()
                exp[ni, i, j] = 0.3 * ((a @ gg) @ Wo) + X[ni, i, j]
    assert np.abs(got - exp).max() < 1e-12


def test_conditional_batchnorm_and_resnet_up_block_semantics():
    """CBN = batch-normalise (biased variance, eps 1e-3, no scale / centre) then * gamma(z) + beta(z) with bias-free Dense
    layers and NO '1 + gamma' (resnet_ops.py:13-28); the up block adds a 1x1 stride-2 transposed-conv shortcut of the RAW
    input whose kernel only reaches even positions while its bias is added everywhere (resnet_ops.py:57-73)."""
    g = torch.Generator().manual_seed(32)
    n, h, w, c, zdim = 3, 2, 3, 4, 5
    x = torch.randn(n, h, w, c, generator=g, dtype=DT) * 2 + 1
    z = torch.randn(n, zdim, generator=g, dtype=DT)
    p = {"c.gamma.w": torch.randn(zdim, c, generator=g, dtype=DT), "c.beta.w": torch.randn(zdim, c, generator=g, dtype=DT),
         "c.moving_mean": torch.zeros(c, dtype=DT), "c.moving_var": torch.ones(c, dtype=DT)}
    stats = {}
    got = O.conditional_batchnorm(x, z, p, "c", True, stats).numpy()
    X = x.numpy()
    mean, var = X.mean(axis=(0, 1, 2)), X.var(axis=(0, 1, 2))
    gam, bet = z.numpy() @ p["c.gamma.w"].numpy(), z.numpy() @ p["c.beta.w"].numpy()
    exp = (X - mean) / np.sqrt(var + 1e-3) * gam[:, None, None, :] + bet[:, None, None, :]
    assert np.abs(got - exp).max() < 1e-12
    cnt = n * h * w
    assert np.allclose(stats["c.moving_var"].numpy(), 0.99 + 0.01 * var * cnt / (cnt - 1))
    # shortcut of the up block
    ws = torch.randn(1, 1, 6, c, generator=g, dtype=DT)
    bs = torch.randn(6, generator=g, dtype=DT)
    s = O.conv2d_transpose(x, ws, bs, (2, 2)).numpy()
    val = X @ ws[0, 0].numpy().T
    assert np.allclose(s[:, ::2, ::2], val + bs.numpy()) and np.allclose(s[:, 1::2, :], bs.numpy()) and np.allclose(s[:, :, 1::2], bs.numpy())
    s21 = O.conv2d_transpose(x, ws, bs, (2, 1)).numpy()
    assert s21.shape == (n, 2 * h, w, 6) and np.allclose(s21[:, ::2], val + bs.numpy()) and np.allclose(s21[:, 1::2], bs.numpy())


def test_resnet_down_block_relu_on_raw_input_and_pool_commutes():
    """ResNetBlockDown (resnet_ops.py:93-115): ReLU is applied to the block INPUT before conv1 (even to raw images, Q11), the
    1x1 shortcut sees the un-rectified input, and avgpool(a) + avgpool(b) == avgpool(a + b) (what the CUDA path exploits)."""
    g = torch.Generator().manual_seed(33)
    n, h, w, ci, co = 2, 4, 4, 3, 5
    x = torch.randn(n, h, w, ci, generator=g, dtype=DT)
    p = {"b.conv1.w": torch.randn(3, 3, ci, co, generator=g, dtype=DT), "b.conv1.b": torch.randn(co, generator=g, dtype=DT),
         "b.conv2.w": torch.randn(3, 3, co, co, generator=g, dtype=DT), "b.conv2.b": torch.randn(co, generator=g, dtype=DT),
         "b.short.w": torch.randn(1, 1, ci, co, generator=g, dtype=DT), "b.short.b": torch.randn(co, generator=g, dtype=DT)}
    got = O.resnet_block_down(x, p, "b", False)
    main = O.conv2d(torch.relu(O.conv2d(torch.relu(x), p["b.conv1.w"], p["b.conv1.b"])), p["b.conv2.w"], p["b.conv2.b"])
    short = O.conv2d(x, p["b.short.w"], p["b.short.b"])
    summed = (main + short).numpy().reshape(n, h // 2, 2, w // 2, 2, co).mean(axis=(2, 4))
    assert np.abs(got.numpy() - summed).max() < 1e-12
    neg = -torch.rand(n, h, w, ci, generator=g, dtype=DT)             # an all-negative input only survives through the shortcut
    only_short = O.resnet_block_down(neg, p, "b", True)
    bias_path = O.conv2d(torch.relu(O.conv2d(torch.zeros_like(neg), p["b.conv1.w"], p["b.conv1.b"])), p["b.conv2.w"], p["b.conv2.b"])
    assert np.abs((only_short - bias_path - O.conv2d(neg, p["b.short.w"], p["b.short.b"])).numpy()).max() < 1e-12


def test_ctc_gradient_matches_finite_differences():
    """d loss / d probs of K.ctc_batch_cost's restatement (eps + re-softmax inside) against central differences in fp64."""
    g = torch.Generator().manual_seed(34)
    t, c = 7, 5
    probs = torch.softmax(torch.randn(1, t, c, generator=g, dtype=DT), -1).requires_grad_(True)
    labels = torch.tensor([[1, 1, 3]])
    il, ll = torch.tensor([[t]]), torch.tensor([[3]])
    loss = O.ctc_batch_cost(labels, probs, il, ll).sum()
    (grad,) = torch.autograd.grad(loss, probs)
    base = probs.detach()
    eps = 1e-6
    for (ti, ci) in ((0, 1), (3, 4), (6, 3), (2, 0)):
        pp, pm = base.clone(), base.clone()
        pp[0, ti, ci] += eps
        pm[0, ti, ci] -= eps
        fd = (O.ctc_batch_cost(labels, pp, il, ll).sum() - O.ctc_batch_cost(labels, pm, il, ll).sum()) / (2 * eps)
        assert abs(float(fd) - float(grad[0, ti, ci])) <= 1e-6 * max(1.0, abs(float(fd)))


def test_losses_match_numpy_formulas():
    """hinge / not_saturating (net_loss.py:4-54) against the literal numpy formulas, incl. the positional quirk of
    not_saturating (3rd arg: style images -> label 1, 4th: training images -> label 0)."""
    g = torch.Generator().manual_seed(35)
    a, b_, c_, d_, e_ = (torch.randn(6, 1, generator=g, dtype=DT) for _ in range(5))
    A, B, C, D_, E = (t.numpy() for t in (a, b_, c_, d_, e_))
    out = [t.numpy() for t in O.hinge(a, b_, c_, d_)]
    relu = lambda v: np.maximum(v, 0)
    exp = [relu(1 - A) + relu(1 + B), relu(1 - A), relu(1 + B), -(B + D_), relu(1 - C) + relu(1 + D_), relu(1 - C), relu(1 + D_)]
    assert all(np.abs(o - e).max() < 1e-12 for o, e in zip(out, exp))
    sce = lambda x, z: np.maximum(x, 0) - x * z + np.log1p(np.exp(-np.abs(x)))
    out = [t.numpy() for t in O.not_saturating(a, b_, c_, d_, e_)]
    exp = [sce(A, 1) + sce(B, 0), sce(A, 1), sce(B, 0), sce(B, 1) + sce(E, 1), sce(C, 1) + sce(D_, 0), sce(C, 1), sce(D_, 0)]
    assert all(np.abs(o - e).max() < 1e-12 for o, e in zip(out, exp))
