"""GPU parity tests of every libsgan primitive, called through the C ABI, against the CPU oracle
(oracle/sgan_oracle.py, fp64) on the same seeded inputs.  Tolerances are written next to each check:
exact-fp32 kernels 1e-4 relative (to the max magnitude of the expected tensor); tensor-core kernels are compared
on inputs already rounded to the operand precision, so only accumulation order differs (2e-3)."""
import importlib

import pytest
import torch

import sgan_oracle as O

pytestmark = pytest.mark.gpu

ops = importlib.import_module("scrabble-gan_b200.ops")
abi = importlib.import_module("scrabble-gan_b200._abi")
F32, BF16 = abi.SG_F32, abi.SG_BF16


def rel_err(got, exp):
    got = got.detach().double().cpu()
    exp = exp.detach().double().cpu()
    assert got.shape == exp.shape, (got.shape, exp.shape)
    assert torch.isfinite(got).all(), "non-finite values in result"
    return float((got - exp).abs().max() / (exp.abs().max() + 1e-30))


def check(got, exp, tol, what=""):
    e = rel_err(got, exp)
    assert e <= tol, "{}: rel err {:.3e} > {:.1e}".format(what, e, tol)


def rnd(gen, *shape):
    return torch.randn(*shape, generator=gen, dtype=torch.float64)


def dev(rt, t, dt=F32):
    return t.to(device=rt.device, dtype=torch.float32 if dt == F32 else torch.bfloat16).contiguous()


def quant(t, mode):
    """Round to the operand precision of a tensor-core mode (bf16: RN to 8 bits; tf32: RN to 11 bits)."""
    if mode == "bf16":
        return t.float().bfloat16().double()
    if mode == "tf32":
        f = t.float().contiguous()
        i = f.view(torch.int32)
        i = ((i + 0x0FFF + ((i >> 13) & 1)) & ~0x1FFF)
        return i.view(torch.float32).double()
    return t.float().double()


# ----------------------------------------------------------------------------------------------------
# convolutions
# ----------------------------------------------------------------------------------------------------
CONV_CASES_SIMT = [
    # n, h, w, ci, co, k, padding
    (2, 8, 12, 1, 64, 3, "same"),
    (2, 6, 10, 64, 1, 3, "same"),
    (2, 8, 8, 64, 8, 1, "same"),
    (3, 4, 10, 32, 64, 1, "same"),
    (2, 2, 9, 16, 24, 2, "valid"),
    (1, 5, 7, 3, 5, 3, "same"),
    (2, 8, 12, 1, 64, 1, "same"),      # Cin = 1 shortcut (1 tap)
    (3, 5, 7, 1, 64, 3, "same"),       # pixel count not a multiple of the 64-pixel chunk
    (3, 5, 7, 64, 1, 3, "same"),
    (2, 4, 9, 1, 32, 3, "same"),       # Cin = 1, 4 channel groups
    (2, 8, 16, 64, 1, 3, "same"),      # Cout = 1, 4 pixels x 8 channels per thread (width % 4 == 0)
    (3, 4, 4, 64, 1, 1, "same"),       # ... single tap, one quad per row
    (3, 4, 4, 1, 64, 3, "same"),
]


@pytest.mark.parametrize("case", CONV_CASES_SIMT)
def test_conv_simt_fwd_dgrad_wgrad(rt, case):
    n, h, w, ci, co, k, padding = case
    g = torch.Generator().manual_seed(1)
    x = rnd(g, n, h, w, ci).requires_grad_(True)
    wt = (rnd(g, k, k, ci, co) * 0.2).requires_grad_(True)
    b = rnd(g, co).requires_grad_(True)
    y = O.conv2d(x, wt, b, padding)
    dy = rnd(g, *y.shape)
    y.backward(dy)

    d = ops.desc_conv_fwd(n, h, w, ci, co, k, k, padding)
    out = rt.empty(y.shape)
    ops.conv_run(rt, d, dev(rt, x), dev(rt, wt), None, dev(rt, b), None, out)
    check(out, y, 1e-4, "fwd")

    # relu + mask + accumulate epilogue
    d2 = ops.desc_conv_fwd(n, h, w, ci, co, k, k, padding, relu=1, accumulate=1)
    m = rnd(g, *y.shape)
    out2 = dev(rt, torch.ones_like(y))
    ops.conv_run(rt, d2, dev(rt, x), dev(rt, wt), None, dev(rt, b), dev(rt, m), out2)
    check(out2, 1 + torch.relu(y) * (m > 0), 1e-4, "fwd epilogue")

    dd = ops.desc_conv_dgrad(n, h, w, ci, co, k, k, padding)
    dx = rt.empty(x.shape)
    ops.conv_run(rt, dd, dev(rt, dy), dev(rt, wt), None, None, None, dx)
    check(dx, x.grad, 1e-4, "dgrad")

    dw = rt.zeros(wt.shape)
    ops.conv_wgrad(rt, d, dev(rt, x), dev(rt, dy), dw, force_simt=True)
    check(dw, wt.grad, 1e-4, "wgrad")

    db = rt.zeros((co,))
    ops.colsum_into(rt, dev(rt, dy), co, db)
    check(db, b.grad, 1e-4, "dbias")


CONV_CASES_TC = [
    # n, h, w, ci, co, k, padding
    (4, 8, 20, 64, 64, 3, "same"),
    (3, 4, 10, 128, 256, 3, "same"),      # multi-image boxes
    (2, 16, 40, 64, 512, 3, "same"),
    (5, 8, 14, 64, 128, 3, "same"),       # ragged width (L = 7)
    (2, 4, 10, 1024, 1024, 3, "same"),    # K = 9216
    (3, 8, 20, 512, 1024, 1, "same"),     # 1x1 shortcut
    (2, 2, 19, 512, 512, 2, "valid"),     # R.conv7
    (1, 32, 160, 64, 64, 3, "same"),      # wide image, rows split across tiles
    (7, 4, 2, 256, 128, 3, "same"),       # L = 1 deepest block
]


@pytest.mark.parametrize("mode", ["bf16", "tf32"])
@pytest.mark.parametrize("case", CONV_CASES_TC)
def test_conv_tc_fwd_dgrad_wgrad(rt, case, mode):
    n, h, w, ci, co, k, padding = case
    rt.set_mode(mode)
    try:
        dt = rt.op_dt
        g = torch.Generator().manual_seed(2)
        x = quant(rnd(g, n, h, w, ci), mode).requires_grad_(True)
        wt = quant(rnd(g, k, k, ci, co) * (1.0 / (k * k * ci) ** 0.5), mode).requires_grad_(True)
        b = rnd(g, co)
        y = O.conv2d(x, wt, b, padding)
        dy = quant(rnd(g, *y.shape), mode)
        y.backward(dy)
        xd, wd, dyd = dev(rt, x, dt), dev(rt, wt), dev(rt, dy, dt)

        d = ops.desc_conv_fwd(n, h, w, ci, co, k, k, padding, in_dt=dt)
        assert ops.tc_ok(rt, d)
        wp = ops.pack_weights(rt, d, wd)
        out = rt.empty(y.shape)
        ops.conv_run(rt, d, xd, wd, wp, dev(rt, b), None, out)
        check(out, y, 2e-3, "tc fwd")

        # bf16/relu/mask epilogue + accumulate variant
        d2 = ops.desc_conv_fwd(n, h, w, ci, co, k, k, padding, in_dt=dt, out_dt=dt, relu=1, mask_dt=dt)
        m = rnd(g, *y.shape)
        out2 = rt.empty(y.shape, dt)
        ops.conv_run(rt, d2, xd, wd, wp, dev(rt, b), dev(rt, m, dt), out2)
        check(out2, torch.relu(y) * (m > 0), 1e-2 if mode == "bf16" else 2e-3, "tc fwd epilogue")
        d3 = ops.desc_conv_fwd(n, h, w, ci, co, k, k, padding, in_dt=dt, accumulate=1)
        out3 = dev(rt, torch.ones_like(y))
        ops.conv_run(rt, d3, xd, wd, wp, None, None, out3)
        check(out3, 1 + y - b, 2e-3, "tc fwd accumulate")

        dd = ops.desc_conv_dgrad(n, h, w, ci, co, k, k, padding, in_dt=dt)
        assert ops.tc_ok(rt, dd)
        wpd = ops.pack_weights(rt, dd, wd)
        dx = rt.empty(x.shape)
        ops.conv_run(rt, dd, dyd, wd, wpd, None, None, dx)
        check(dx, x.grad, 2e-3, "tc dgrad")

        dw = rt.zeros(wt.shape)
        dwg = ops.desc_conv_fwd(n, h, w, ci, co, k, k, padding, in_dt=dt, out_dt=dt)
        ops.conv_wgrad(rt, dwg, xd, dyd, dw)
        check(dw, wt.grad, 2e-3, "tc wgrad")
        # filter gradient + bias gradient(s) in one launch (dy^T . 1 on the tensor cores), accumulating into both biases
        dw2, db, db2 = rt.zeros(wt.shape), rt.zeros((co,)), torch.ones(co, device=rt.device)
        assert ops.conv_wgrad(rt, dwg, xd, dyd, dw2, db=db, db2=db2, force_bias=True), "a plain Conv2D on the tensor-core path can fuse its bias gradient"
        assert torch.equal(dw2, dw), "the bias units must not change the filter gradient"
        exp_db = dy.sum(dim=(0, 1, 2))
        check(db, exp_db, 2e-3, "tc wgrad fused bias gradient")
        check(db2 - 1.0, exp_db, 2e-3, "tc wgrad fused second bias gradient")

        if mode == "bf16":
            # pack-free launches: the filter is read in place from a bf16 mirror of the HWIO master (N-major B operand for
            # the forward conv, K-major for the dgrad)
            wm = wd.to(torch.bfloat16)
            if ops.direct_ok(rt, d, force=True):
                out4 = rt.empty(y.shape)
                ops.conv_run(rt, d, xd, wd, None, dev(rt, b), None, out4, w_mirror=wm)
                check(out4, y, 2e-3, "tc fwd (direct weights)")
                out5 = rt.empty(y.shape, dt)
                ops.conv_run(rt, d2, xd, wd, None, dev(rt, b), dev(rt, m, dt), out5, w_mirror=wm)
                check(out5, torch.relu(y) * (m > 0), 1e-2, "tc fwd epilogue (direct weights)")
            else:
                assert co % 64 != 0, "HWIO forward convs with c_out % 64 == 0 must support the direct path"
            assert ops.direct_ok(rt, dd)
            dx2 = rt.empty(x.shape)
            ops.conv_run(rt, dd, dyd, wd, None, None, None, dx2, w_mirror=wm)
            check(dx2, x.grad, 2e-3, "tc dgrad (direct weights)")
    finally:
        rt.set_mode("fp32")


@pytest.mark.parametrize("case", [(128, 4, 10, 1024, 1024, 3), (64, 4, 10, 1024, 1024, 3), (64, 8, 20, 1024, 512, 3), (2, 4, 10, 1024, 1024, 3),
                                  (128, 4, 10, 1024, 1024, 1)])
def test_conv_tc_split_tail(rt, case):
    """Wave quantisation: the tiles of a partial last wave are split along k over several CTAs (sg_ctx_set_conv_split_tail).
    Same result as the unsplit launch up to the fp32 summation order, bit-identical run to run, and the rendezvous counters are
    re-armed (the launch can be repeated)."""
    n, h, w, ci, co, k = case
    rt.set_mode("bf16")
    try:
        g = torch.Generator().manual_seed(11)
        x = dev(rt, rnd(g, n, h, w, ci), BF16)
        wt = dev(rt, rnd(g, k, k, ci, co) * (1.0 / (k * k * ci) ** 0.5))
        b = dev(rt, rnd(g, co))
        m = dev(rt, rnd(g, n, h, w, co), BF16)
        outs = {}
        for split in (0, 1):
            ops.call.sg_ctx_set_conv_split_tail(rt.ctx, split)
            res = []
            for rep in range(3):
                d = ops.desc_conv_fwd(n, h, w, ci, co, k, k, "same", in_dt=BF16)
                wp = ops.pack_weights(rt, d, wt)
                out = rt.empty((n, h, w, co))
                ops.conv_run(rt, d, x, wt, wp, b, None, out)
                d2 = ops.desc_conv_fwd(n, h, w, ci, co, k, k, "same", in_dt=BF16, out_dt=BF16, relu=1, mask_dt=BF16)
                out2 = rt.empty((n, h, w, co), BF16)
                ops.conv_run(rt, d2, x, wt, wp, b, m, out2)
                dd = ops.desc_conv_dgrad(n, h, w, co, ci, k, k, "same", in_dt=BF16) if ci == co else None
                out3 = None
                if dd is not None:
                    out3 = rt.empty((n, h, w, ci))
                    ops.conv_run(rt, dd, x, wt, None, None, None, out3, w_mirror=wt.to(torch.bfloat16))
                res.append((out, out2, out3))
            for r in res[1:]:
                for a, b_ in zip(res[0], r):
                    assert a is None or torch.equal(a, b_), "split-tail launches must be bit-reproducible"
            outs[split] = res[0]
        for a, b_ in zip(outs[0], outs[1]):
            if a is not None:
                assert rel_err(b_, a) <= (2e-5 if a.dtype == torch.float32 else 1e-2), "split and unsplit launches differ"
    finally:
        ops.call.sg_ctx_set_conv_split_tail(rt.ctx, 1)
        rt.set_mode("fp32")


@pytest.mark.parametrize("mode", ["bf16", "tf32"])
@pytest.mark.parametrize("case", [(3, 8, 20, 128, 256, 64), (2, 4, 10, 256, 512, 512), (4, 16, 40, 64, 64, 64)])
def test_conv_tc_with_fused_shortcut(rt, case, mode):
    """conv3x3(h1) + conv1x1(xs) + bias in ONE tensor-core launch (ResNetBlockDown: resnet_ops.py:103-114)."""
    n, h, w, ci, co, ci2 = case
    rt.set_mode(mode)
    try:
        dt = rt.op_dt
        g = torch.Generator().manual_seed(17)
        x = quant(rnd(g, n, h, w, ci), mode)
        x2 = quant(rnd(g, n, h, w, ci2), mode)
        w1 = quant(rnd(g, 3, 3, ci, co) * (1.0 / (9 * ci) ** 0.5), mode)
        w2 = quant(rnd(g, 1, 1, ci2, co) * (1.0 / ci2 ** 0.5), mode)
        b = rnd(g, co)
        y = O.conv2d(x, w1, None, "same") + O.conv2d(x2, w2, None, "same") + b
        d = ops.desc_conv_fwd(n, h, w, ci, co, 3, 3, "same", in_dt=dt)
        d2 = ops.desc_conv_fwd(n, h, w, ci2, co, 1, 1, "same", in_dt=dt)
        w1d, w2d = dev(rt, w1), dev(rt, w2)
        out = rt.empty(y.shape)
        ops.conv_run_dual(rt, d, dev(rt, x, dt), ops.pack_weights(rt, d, w1d), d2, dev(rt, x2, dt), ops.pack_weights(rt, d2, w2d),
                          dev(rt, b), None, out)
        check(out, y, 2e-3, "conv + fused 1x1 shortcut")
    finally:
        rt.set_mode("fp32")


CONVT_CASES = [
    # n, h, w, ci, co, k, sy, sx
    (2, 4, 8, 128, 64, 3, 2, 2),
    (3, 4, 20, 512, 256, 3, 2, 2),
    (2, 16, 20, 128, 64, 3, 2, 1),
    (2, 4, 8, 128, 64, 1, 2, 2),
    (2, 8, 12, 64, 64, 1, 2, 1),
]


@pytest.mark.parametrize("mode", ["fp32", "bf16", "tf32"])
@pytest.mark.parametrize("case", CONVT_CASES)
def test_conv_transpose(rt, case, mode):
    n, h, w, ci, co, k, sy, sx = case
    rt.set_mode(mode)
    try:
        dt = rt.op_dt
        tol = 1e-4 if mode == "fp32" else 2e-3
        g = torch.Generator().manual_seed(3)
        x = quant(rnd(g, n, h, w, ci), mode).requires_grad_(True)
        wt = quant(rnd(g, k, k, co, ci) * (1.0 / (k * k * ci) ** 0.5), mode).requires_grad_(True)
        b = rnd(g, co)
        y = O.conv2d_transpose(x, wt, b, (sy, sx))
        dy = quant(rnd(g, *y.shape), mode)
        y.backward(dy)
        xd, wd, dyd = dev(rt, x, dt), dev(rt, wt), dev(rt, dy, dt)

        # forward: bias everywhere, then one accumulate launch per phase that has taps
        out = dev(rt, (b.view(1, 1, 1, co)).expand(n, h * sy, w * sx, co).clone())
        for (py, px) in ops.convT_phases(k, sy, sx):
            d = ops.desc_convT_phase(n, h, w, ci, co, k, sy, sx, py, px, in_dt=dt, accumulate=1)
            wp = ops.pack_weights(rt, d, wd) if ops.tc_ok(rt, d) else None
            ops.conv_run(rt, d, xd, wd, wp, None, None, out)
        check(out, y, tol, "convT fwd")

        dd = ops.desc_convT_dgrad(n, h, w, ci, co, k, sy, sx, in_dt=dt)
        wpd = ops.pack_weights(rt, dd, wd) if ops.tc_ok(rt, dd) else None
        dx = rt.empty(x.shape)
        ops.conv_run(rt, dd, dyd, wd, wpd, None, None, dx)
        check(dx, x.grad, tol, "convT dgrad")

        dw = rt.zeros(wt.shape)
        dwg = ops.desc_convT_dgrad(n, h, w, ci, co, k, sy, sx, in_dt=dt, out_dt=dt)
        ops.conv_wgrad(rt, dwg, dyd, xd, dw)
        check(dw, wt.grad, tol, "convT wgrad")

        if mode == "bf16":       # pack-free launches on the (kh,kw,Cout,Cin) master: phases K-major, dgrad N-major
            wm = wd.to(torch.bfloat16)
            out2 = dev(rt, (b.view(1, 1, 1, co)).expand(n, h * sy, w * sx, co).clone())
            for (py, px) in ops.convT_phases(k, sy, sx):
                d = ops.desc_convT_phase(n, h, w, ci, co, k, sy, sx, py, px, in_dt=dt, accumulate=1)
                assert ops.direct_ok(rt, d)
                ops.conv_run(rt, d, xd, wd, None, None, None, out2, w_mirror=wm)
            check(out2, y, tol, "convT fwd (direct weights)")
            if ops.direct_ok(rt, dd, force=True):
                dx2 = rt.empty(x.shape)
                ops.conv_run(rt, dd, dyd, wd, None, None, None, dx2, w_mirror=wm)
                check(dx2, x.grad, tol, "convT dgrad (direct weights)")
    finally:
        rt.set_mode("fp32")


# ----------------------------------------------------------------------------------------------------
# element-wise, pooling
# ----------------------------------------------------------------------------------------------------
def test_elementwise(rt):
    g = torch.Generator().manual_seed(4)
    x = rnd(g, 3, 5, 7, 12)
    xd = dev(rt, x)
    r, c = ops.act_prep(rt, xd, True, True, BF16)
    check(r, torch.relu(x), 1e-2, "relu bf16")
    check(c, x, 1e-2, "copy bf16")
    r, _ = ops.act_prep(rt, xd, True, False, F32)
    assert torch.equal(r.cpu().double(), torch.relu(x.float()).double())
    dy = rnd(g, *x.shape)
    out = ops.mask_mul(rt, dev(rt, dy), r, F32)
    assert torch.equal(out.cpu(), (dy.float() * (x.float() > 0)))
    out = ops.mask_mul(rt, dev(rt, dy), r, F32, out=out, accumulate=1)
    check(out, 2 * dy * (x > 0), 1e-6, "mask_mul accumulate")
    y = rnd(g, *x.shape)
    check(ops.axpby(rt, 0.5, xd, -2.0, dev(rt, y)), 0.5 * x - 2 * y, 1e-6, "axpby")
    sig = dev(rt, torch.tensor([0.37], dtype=torch.float64))
    check(ops.scale_add(rt, sig, xd, dev(rt, y)), 0.37 * x + y, 1e-6, "scale_add")
    check(ops.scale_add(rt, sig, xd, None), 0.37 * x, 1e-6, "scale")
    t = ops.tanh_fwd(rt, xd)
    check(t, torch.tanh(x), 1e-5, "tanh")
    check(ops.tanh_bwd(rt, dev(rt, dy), t), dy * (1 - torch.tanh(x) ** 2), 1e-5, "tanh bwd")
    w = rnd(g, 3)
    z = dev(rt, x.clone())
    ops.scale_rows_(rt, z, dev(rt, w))
    check(z, x * w.view(3, 1, 1, 1), 1e-6, "scale_rows")
    acc = rt.zeros((1,))
    ops.dot_into(rt, xd, dev(rt, y), acc, accumulate=0)
    check(acc, (x * y).sum().view(1), 1e-5, "dot")


@pytest.mark.parametrize("rows,cols", [(5000, 64), (777, 1024), (1283, 512), (9, 8), (300, 53), (4096, 1), (70001, 32), (33, 2048)])
@pytest.mark.parametrize("dt", [F32, BF16])
def test_colsum(rt, rows, cols, dt):
    """Bias gradients: out[c] (+)= sum_r x[r,c]; vectorised path (cols % 4 == 0) and the generic fallback."""
    g = torch.Generator().manual_seed(44)
    x = rnd(g, rows, cols)
    if dt == BF16:
        x = x.float().bfloat16().double()
    out = rt.zeros((cols,))
    ops.colsum_into(rt, dev(rt, x, dt), cols, out, accumulate=0)
    exp = x.sum(0)
    assert float((out.cpu().double() - exp).abs().max()) <= 1e-5 * float(x.abs().sum(0).max()), "colsum"
    ops.colsum_into(rt, dev(rt, x, dt), cols, out, accumulate=1)
    assert float((out.cpu().double() - 2 * exp).abs().max()) <= 2e-5 * float(x.abs().sum(0).max()), "colsum accumulate"


def test_pooling(rt):
    g = torch.Generator().manual_seed(5)
    x = rnd(g, 2, 8, 12, 16).requires_grad_(True)
    y = O.avg_pool_2x2_same(x)
    dy = rnd(g, *y.shape)
    y.backward(dy)
    check(ops.avgpool2_fwd(rt, dev(rt, x)), y, 1e-6, "avgpool fwd")
    check(ops.avgpool2_bwd(rt, dev(rt, dy), F32), x.grad, 1e-6, "avgpool bwd")
    for ph, pw in ((2, 2), (2, 1)):
        x = rnd(g, 2, 8, 12, 16).requires_grad_(True)
        a = torch.relu(x)
        y = O.max_pool(a, ph, pw)
        dy = rnd(g, *y.shape)
        y.backward(dy)
        ad = dev(rt, a)
        check(ops.maxpool_fwd(rt, ad, ph, pw), y, 1e-6, "maxpool fwd")
        check(ops.maxpool_bwd(rt, dev(rt, dy), ad, ph, pw, True, F32), x.grad, 1e-6, "maxpool bwd (relu gated)")
    x = rnd(g, 3, 4, 10, 256).requires_grad_(True)
    y = torch.relu(x).mean(dim=(1, 2))
    dy = rnd(g, *y.shape)
    y.backward(dy)
    check(ops.gap_relu_fwd(rt, dev(rt, x)), y, 1e-5, "gap fwd")
    check(ops.gap_relu_bwd(rt, dev(rt, dy), dev(rt, x)), x.grad, 1e-6, "gap bwd")


# ----------------------------------------------------------------------------------------------------
# batch norm
# ----------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("c,per_sample", [(64, True), (512, True), (64, False), (256, True)])
def test_batchnorm_train_fwd_bwd(rt, c, per_sample):
    g = torch.Generator().manual_seed(6)
    n, h, w = 3, 4, 6
    x = (rnd(g, n, h, w, c) * 1.7 + 0.4).requires_grad_(True)
    gamma = (rnd(g, n, c) if per_sample else rnd(g, c)).requires_grad_(True)
    beta = (rnd(g, n, c) if per_sample else rnd(g, c)).requires_grad_(True)
    mm, mv = rnd(g, c) * 0.1, rnd(g, c).abs() + 0.5
    xh, nm, nv = O.batchnorm_train(x, mm, mv)
    gb = (lambda t: t.view(n, 1, 1, c)) if per_sample else (lambda t: t)
    y = torch.relu(xh * gb(gamma) + gb(beta))
    dy = rnd(g, *y.shape)
    y.backward(dy)

    xd = dev(rt, x)
    sums = ops.bn_stats(rt, xd)
    mmd, mvd = dev(rt, mm), dev(rt, mv)
    mean, rstd = ops.bn_finalize(rt, sums, n * h * w, c, mmd, mvd)
    check(mmd, nm, 1e-5, "moving mean")
    check(mvd, nv, 1e-5, "moving var")
    gd, bd = dev(rt, gamma), dev(rt, beta)
    act = ops.bn_apply(rt, xd, mean, rstd, gd, bd, per_sample, True, F32)
    check(act, y, 1e-5, "bn apply")
    s1, s2 = ops.bn_bwd_reduce(rt, dev(rt, dy), act, xd, mean, rstd)
    ab = ops.bn_bwd_combine(rt, s1, s2, gd, per_sample)
    dx = ops.bn_bwd_apply(rt, dev(rt, dy), act, xd, mean, rstd, gd, per_sample, ab, n * h * w, True, False, F32)
    check(dx, x.grad, 1e-4, "bn dx")
    if per_sample:
        check(s2, gamma.grad, 1e-4, "dgamma")
        check(s1, beta.grad, 1e-4, "dbeta")
    else:
        check(s2.sum(0), gamma.grad, 1e-4, "dgamma")
        check(s1.sum(0), beta.grad, 1e-4, "dbeta")


def test_batchnorm_inference_bwd(rt):
    """R's BN: x = relu(conv) -> inference-mode BN with gamma/beta; dx gated by x > 0."""
    g = torch.Generator().manual_seed(7)
    n, h, w, c = 2, 4, 9, 512
    pre = rnd(g, n, h, w, c).requires_grad_(True)
    gamma, beta = rnd(g, c).requires_grad_(True), rnd(g, c).requires_grad_(True)
    mm, mv = rnd(g, c) * 0.1, rnd(g, c).abs() + 0.5
    x = torch.relu(pre)
    y = O.batchnorm_infer(x, mm, mv) * gamma + beta
    dy = rnd(g, *y.shape)
    y.backward(dy)
    xd = dev(rt, x)
    mean, rstd = ops.bn_infer_prepare(rt, dev(rt, mm), dev(rt, mv))
    out = ops.bn_apply(rt, xd, mean, rstd, dev(rt, gamma), dev(rt, beta), False, False, BF16)
    check(out, y, 1e-2, "bn infer apply (bf16 out)")
    s1, s2 = ops.bn_bwd_reduce(rt, dev(rt, dy), None, xd, mean, rstd)
    check(s2.sum(0), gamma.grad, 1e-4, "dgamma")
    check(s1.sum(0), beta.grad, 1e-4, "dbeta")
    dx = ops.bn_bwd_apply(rt, dev(rt, dy), None, xd, mean, rstd, dev(rt, gamma), False, None, 1.0, False, True, F32)
    check(dx, pre.grad, 1e-5, "dx gated")


# ----------------------------------------------------------------------------------------------------
# dense / filter bank / attention / CTC / losses / optimizers / spectral norm
# ----------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("m,n,k,ta,tb", [(64, 128, 1024, 0, 0), (7, 81, 512, 0, 0), (64, 8, 20000, 1, 0), (33, 32, 64, 0, 1),
                                         (5, 1, 1024, 0, 0), (1024, 1, 5, 1, 0)])
def test_gemm(rt, m, n, k, ta, tb):
    g = torch.Generator().manual_seed(8)
    a = rnd(g, k, m) if ta else rnd(g, m, k)
    b = rnd(g, n, k) if tb else rnd(g, k, n)
    bias = rnd(g, n)
    exp = (a.t() if ta else a) @ (b.t() if tb else b) + bias
    out = ops.gemm(rt, dev(rt, a), dev(rt, b), m, n, k, bool(ta), bool(tb), bias=dev(rt, bias))
    check(out, exp, 2e-5, "gemm")
    out = ops.gemm(rt, dev(rt, a), dev(rt, b), m, n, k, bool(ta), bool(tb), out=out, accumulate=1)
    check(out, 2 * exp - bias, 2e-5, "gemm accumulate")


def test_filterbank(rt):
    g = torch.Generator().manual_seed(9)
    b, l, vocab = 3, 4, 11
    bank = (rnd(g, vocab, 32, 8192) * 0.05).requires_grad_(True)
    z = rnd(g, b, 128).requires_grad_(True)
    y = torch.randint(0, vocab, (b, l), generator=g)
    y[0, 1] = y[0, 0]        # repeated character
    out = O.filter_bank(z[:, :32], y, bank)
    dout = rnd(g, *out.shape)
    out.backward(dout)
    yd = y.to(rt.device, torch.int32)
    got = ops.filterbank_fwd(rt, dev(rt, z), 128, yd, dev(rt, bank))
    check(got, out, 1e-5, "filterbank fwd")
    # bit-exact character -> filter indexing: a one-hot z picks single bank rows which must land unrounded
    zi = torch.zeros(b, 128, dtype=torch.float64)
    zi[:, 5] = 1.0
    got1 = ops.filterbank_fwd(rt, dev(rt, zi), 128, yd, dev(rt, bank)).cpu()
    bank32 = bank.detach().float()
    for bi in range(b):
        for li in range(l):
            for k in range(0, 8192, 61):
                hh, ww, cc = O.filter_bank_index_map(li, k)
                assert got1[bi, hh, ww, cc].item() == bank32[y[bi, li], 5, k].item()
    dbank = rt.empty(bank.shape)
    dz0 = rt.empty((b, 32))
    ops.filterbank_bwd(rt, dev(rt, dout), dev(rt, z), 128, yd, dev(rt, bank), dbank, dz0, 32)
    check(dbank, bank.grad, 1e-5, "dbank")
    check(dz0, z.grad[:, :32], 1e-5, "dz0")


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
@pytest.mark.parametrize("p", [77, 128 * 3, 5000, 20001])
def test_nonlocal_projections(rt, p, mode):
    """Fused 1x1 projections of the non-local block (arch_ops.py:38-67), forward and backward, vs plain matmuls: exact FFMA
    kernels in fp32 mode (1e-5), bf16 warp-level tensor ops in the speed mode (1e-2 of the largest magnitude)."""
    rt.set_mode(mode)
    try:
        _nonlocal_projections_case(rt, p, 1e-5 if mode == "fp32" else 1e-2)
    finally:
        rt.set_mode("fp32")


def _nonlocal_projections_case(rt, p, tol):
    g = torch.Generator().manual_seed(21)
    x = rnd(g, p, 64).requires_grad_(True)
    wt, wp, wg = [(rnd(g, 64, k) * 0.2).requires_grad_(True) for k in (8, 8, 32)]
    wo = (rnd(g, 32, 64) * 0.2).requires_grad_(True)
    sigma = torch.tensor([0.37], dtype=torch.float64, requires_grad=True)
    theta, phi, gg = x @ wt, x @ wp, x @ wg
    th_d, ph_d, g_d = ops.nonlocal_proj_fwd(rt, dev(rt, x), dev(rt, wt), dev(rt, wp), dev(rt, wg))
    check(th_d, theta, tol, "theta"); check(ph_d, phi, tol, "phi"); check(g_d, gg, tol, "g")
    # projection backward
    dth, dph, dgg = rnd(g, p, 8), rnd(g, p, 8), rnd(g, p, 32)
    dx0 = rnd(g, p, 64)
    (theta * dth).sum().backward(retain_graph=True); (phi * dph).sum().backward(retain_graph=True); (gg * dgg).sum().backward()
    dx = dev(rt, dx0)
    gt, gp, gg_ = rt.zeros((64, 8)), rt.zeros((64, 8)), rt.zeros((64, 32))
    ops.nonlocal_proj_bwd(rt, dev(rt, x), dev(rt, dth), dev(rt, dph), dev(rt, dgg), dev(rt, wt), dev(rt, wp), dev(rt, wg), dx, gt, gp, gg_)
    check(dx, dx0 + x.grad, tol, "dx")
    check(gt, wt.grad, tol, "dw_theta"); check(gp, wp.grad, tol, "dw_phi"); check(gg_, wg.grad, tol, "dw_g")
    dx2 = dev(rt, dx0)
    ops.nonlocal_proj_bwd(rt, dev(rt, x), dev(rt, dth), dev(rt, dph), dev(rt, dgg), dev(rt, wt), dev(rt, wp), dev(rt, wg), dx2)
    check(dx2, dx0 + x.grad, tol, "dx (no wgrad)")
    # output projection
    o = rnd(g, p, 32).requires_grad_(True)
    xr = rnd(g, p, 64)
    og = o @ wo
    out = sigma * og + xr
    og_d, out_d = ops.nonlocal_out_fwd(rt, dev(rt, o), dev(rt, wo), dev(rt, sigma), dev(rt, xr))
    check(og_d, og, tol, "og"); check(out_d, out, tol, "out")
    dout = rnd(g, p, 64)
    (out * dout).sum().backward()
    dwo = rt.zeros((32, 64))
    d_o = ops.nonlocal_out_bwd(rt, dev(rt, dout), dev(rt, o), dev(rt, wo), dev(rt, sigma), dwo)
    check(d_o, o.grad, tol, "d_o"); check(dwo, wo.grad, tol, "dw_o")


@pytest.mark.parametrize("n,h,w", [(2, 8, 12), (1, 32, 80), (3, 16, 40)])
def test_attention(rt, n, h, w):
    g = torch.Generator().manual_seed(10)
    q, kv = h * w, (h // 2) * (w // 2)
    theta = (rnd(g, n, q, 8) * 0.7).requires_grad_(True)
    phi = (rnd(g, n, kv, 8) * 0.7).requires_grad_(True)
    gg = rnd(g, n, kv, 32).requires_grad_(True)
    attn = torch.softmax(theta @ phi.transpose(1, 2), dim=-1)
    o = attn @ gg
    do = rnd(g, *o.shape)
    o.backward(do)
    td, pd, gd = dev(rt, theta), dev(rt, phi), dev(rt, gg)
    od, lse = ops.attn_fwd(rt, td, pd, gd)
    check(od, o, 1e-4, "attn fwd")
    check(lse, torch.logsumexp(theta @ phi.transpose(1, 2), dim=-1), 1e-5, "lse")
    dt_, dp_, dg_ = ops.attn_bwd(rt, td, pd, gd, od, lse, dev(rt, do))
    check(dt_, theta.grad, 2e-4, "dtheta")
    check(dp_, phi.grad, 2e-4, "dphi")
    check(dg_, gg.grad, 2e-4, "dg")


@pytest.mark.parametrize("n,h,w", [(2, 8, 12), (1, 32, 80), (3, 16, 40), (2, 6, 10), (1, 32, 160)])
def test_attention_tensor_core(rt, n, h, w):
    """Speed-mode attention (tf32 theta.phi^T, bf16 P.g via mma.sync) against the fp64 oracle: bf16 tolerance 1e-2 relative
    to the largest magnitude of each tensor (north_star: 1e-2 in bf16)."""
    g = torch.Generator().manual_seed(10)
    q, kv = h * w, (h // 2) * (w // 2)
    theta = (rnd(g, n, q, 8) * 0.7).requires_grad_(True)
    phi = (rnd(g, n, kv, 8) * 0.7).requires_grad_(True)
    gg = rnd(g, n, kv, 32).requires_grad_(True)
    attn = torch.softmax(theta @ phi.transpose(1, 2), dim=-1)
    o = attn @ gg
    do = rnd(g, *o.shape)
    o.backward(do)
    td, pd, gd = dev(rt, theta), dev(rt, phi), dev(rt, gg)
    assert abi.load().sg_attn_tc_supported(q, kv, 8, 32)
    od, lse = ops.attn_fwd(rt, td, pd, gd, tc=True)
    check(od, o, 1e-2, "attn fwd (tc)")
    check(lse, torch.logsumexp(theta @ phi.transpose(1, 2), dim=-1), 2e-3, "lse (tc)")
    dt_, dp_, dg_ = ops.attn_bwd(rt, td, pd, gd, od, lse, dev(rt, do), tc=True)
    check(dt_, theta.grad, 1e-2, "dtheta (tc)")
    check(dp_, phi.grad, 1e-2, "dphi (tc)")
    check(dg_, gg.grad, 1e-2, "dg (tc)")


@pytest.mark.parametrize("b,l,c", [(4, 5, 53), (3, 10, 81), (2, 1, 53), (5, 3, 7)])
def test_ctc(rt, b, l, c):
    g = torch.Generator().manual_seed(11)
    t = 4 * l - 1
    logits = (rnd(g, b, t, c) * 2.0).requires_grad_(True)
    labels = torch.randint(0, c - 1, (b, l), generator=g)
    if l >= 3:
        labels[0, 1] = labels[0, 0]      # repeated label forces a blank between them
        labels[1, :] = labels[1, 0]      # a word of one repeated letter
    probs = torch.softmax(logits, dim=-1)
    loss = O.ctc_batch_cost(labels, probs, torch.full((b, 1), t), torch.full((b, 1), l))
    loss.sum().backward()
    got, grad = ops.ctc(rt, dev(rt, logits), labels.to(rt.device, torch.int32))
    # north_star: CTC loss within 1e-4 relative
    assert float(((got.cpu().double() - loss.view(-1)).abs() / loss.view(-1).abs()).max()) <= 1e-4
    check(grad, logits.grad, 1e-3, "ctc grad wrt dense pre-activations")


def test_ctc_brute_force(rt):
    g = torch.Generator().manual_seed(12)
    t, c = 5, 4
    logits = rnd(g, 1, t, c)
    labels = torch.tensor([[1, 1]])
    probs = torch.softmax(logits, -1)[0]
    q = (probs + 1e-7) / (1 + c * 1e-7)
    exp = O.ctc_brute_force(q, [1, 1], c - 1)
    got, _ = ops.ctc(rt, dev(rt, logits), labels.to(rt.device, torch.int32))
    assert abs(got.item() - exp) / exp <= 1e-4


@pytest.mark.parametrize("kind,use_w,balance", [("hinge", 0, 1), ("hinge", 1, 1), ("hinge", 0, 0), ("not_saturating", 1, 1),
                                                ("not_saturating", 0, 1)])
def test_losses_and_gradient_balance(rt, kind, use_w, balance):
    g = torch.Generator().manual_seed(13)
    b = 37
    names = ["d_real", "d_fake", "s_real", "s_fake", "s5", "r_fake", "r_real"]
    v = {k: rnd(g, b, 1).requires_grad_(True) for k in names}
    with torch.no_grad():
        v["r_fake"].mul_(0.5).add_(30.0)
        v["r_real"].add_(20.0)
    zero = torch.zeros(b, 1, dtype=torch.float64)
    sr, sf, s5 = (v["s_real"], v["s_fake"], v["s5"]) if use_w else (zero, zero, zero)
    if kind == "hinge":
        d_loss, dlr, dlf, g_loss, s_loss, s1, s2 = O.hinge(v["d_real"], v["d_fake"], sr, sf)
        if not use_w:
            g_loss = -v["d_fake"]
    else:
        d_loss, dlr, dlf, g_loss, s_loss, s1, s2 = O.not_saturating(v["d_real"], v["d_fake"], sr, sf, s5)
        if not use_w:
            g_loss = O._sce(v["d_fake"], True)
    g_bal, r_bal, alpha, r_std, g_std = O.apply_gradient_balancing(v["r_fake"], g_loss, 1.0)
    g_added = g_loss + v["r_fake"]
    g_final = g_bal if balance else g_added
    gd = torch.autograd.grad(d_loss.sum(), [v["d_real"], v["d_fake"]], retain_graph=True)
    gg = torch.autograd.grad(g_final.sum(), [v["d_fake"], v["r_fake"]] + ([v["s_fake"]] if use_w else []), retain_graph=True,
                             allow_unused=True)
    kid = abi.SG_LOSS_HINGE if kind == "hinge" else abi.SG_LOSS_NOT_SATURATING
    D = {k: dev(rt, t.detach().view(-1)) for k, t in v.items()}
    sums = torch.empty(abi.SG_LOSS_NSUMS, device=rt.device, dtype=torch.float64)
    P = ops._p
    abi.call.sg_loss_sums(rt.ctx, kid, use_w, P(D["d_real"]), P(D["d_fake"]), P(D["s_real"]), P(D["s_fake"]), P(D["s5"]),
                          P(D["r_fake"]), P(D["r_real"]), b, P(sums))
    ups = [rt.empty((b,)) for _ in range(8)]
    stats = rt.empty((16,))
    abi.call.sg_loss_finish(rt.ctx, kid, use_w, balance, 1.0, P(D["d_real"]), P(D["d_fake"]), P(D["s_real"]), P(D["s_fake"]),
                            P(D["s5"]), P(D["r_fake"]), b, P(sums), *[P(u) for u in ups], P(stats))
    up_d_real, up_d_fake_d, up_s_real, up_s_fake_w, up_s5, up_d_fake_g, up_s_fake_g, up_r_fake_g = ups
    check(up_d_real, gd[0].view(-1), 1e-5, "d/d d_real")
    check(up_d_fake_d, gd[1].view(-1), 1e-5, "d/d d_fake (D loss)")
    check(up_d_fake_g, gg[0].view(-1), 1e-4, "d/d d_fake (G loss)")
    check(up_r_fake_g, gg[1].view(-1), 1e-4, "d/d r_fake (G loss)")
    if use_w:
        gw = torch.autograd.grad(s_loss.sum(), [v["s_real"], v["s_fake"]], retain_graph=True)
        check(up_s_real, gw[0].view(-1), 1e-5, "d/d s_real")
        check(up_s_fake_w, gw[1].view(-1), 1e-5, "d/d s_fake (W loss)")
        exp_sg = gg[2].view(-1) if gg[2] is not None else torch.zeros(b, dtype=torch.float64)
        check(up_s_fake_g, exp_sg, 1e-4, "d/d s_fake (G loss)")
    exp_stats = [v["r_fake"].mean(), v["r_real"].mean(), r_bal.mean(), g_loss.mean(), g_added.mean(), g_bal.mean(),
                 d_loss.mean(), dlr.mean(), dlf.mean(), g_final.mean(), torch.tensor(1.0), r_std, g_std]
    exp_stats += [s_loss.mean(), s1.mean(), s2.mean()] if use_w else [torch.tensor(0.0)] * 3
    exp_stats = torch.stack([t.detach().double().reshape(()) for t in exp_stats])
    got = stats.cpu().double()
    assert float(((got - exp_stats).abs() / (exp_stats.abs() + 1e-3)).max()) <= 1e-4


def test_adam_rmsprop(rt):
    g = torch.Generator().manual_seed(14)
    n = 100003
    w, gr, m, v = rnd(g, n), rnd(g, n), rnd(g, n) * 0.1, rnd(g, n).abs() * 0.1
    import math
    for step, (b1, b2) in ((1, (0.0, 0.999)), (7, (0.5, 0.999))):
        w1, m1, v1 = O.adam_update(w, gr, m, v, step, 2e-4, b1, b2)
        lr_t = 2e-4 * math.sqrt(1 - b2 ** step) / (1 - b1 ** step)
        wd, md, vd = dev(rt, w), dev(rt, m), dev(rt, v)
        ops.adam_(rt, wd, dev(rt, gr), md, vd, lr_t, b1, b2, 1e-7)
        check(wd, w1, 1e-6, "adam w")
        check(md, m1, 1e-6, "adam m")
        check(vd, v1, 1e-6, "adam v")
        # the train step's launch: same update with the step size read from device memory and a bf16 mirror written in the
        # same pass; with beta1 == 0 (m_t = g_t) the first-moment slot is left alone; clear_grad zeroes the consumed gradient
        for clear in (0, 1):
            wd, md, vd, gd = dev(rt, w), dev(rt, m), dev(rt, v), dev(rt, gr)
            mirror = torch.empty(n, device=rt.device, dtype=torch.bfloat16)
            lr_dev = torch.tensor([lr_t], device=rt.device, dtype=torch.float32)
            ops.call.sg_adam_fused(rt.ctx, ops._p(wd), ops._p(gd), ops._p(md), ops._p(vd), ops._p(mirror), n, ops._p(lr_dev), b1, b2, 1e-7, clear)
            check(wd, w1, 1e-6, "fused adam w")
            check(vd, v1, 1e-6, "fused adam v")
            check(md, m if b1 == 0.0 else m1, 1e-6, "fused adam m")
            assert torch.equal(mirror, wd.to(torch.bfloat16))
            assert torch.equal(gd, torch.zeros_like(gd) if clear else dev(rt, gr))
    w1, ms1 = O.rmsprop_update(w, gr, v, 2e-4)
    wd, msd = dev(rt, w), dev(rt, v)
    ops.rmsprop_(rt, wd, dev(rt, gr), msd, 2e-4, 0.9, 1e-7)
    check(wd, w1, 1e-6, "rmsprop w")
    check(msd, ms1, 1e-6, "rmsprop ms")


@pytest.mark.parametrize("shape", [(3, 3, 64, 128), (32, 512), (1, 1, 1024, 1024)])
def test_spectral_norm(rt, shape):
    g = torch.Generator().manual_seed(15)
    w = rnd(g, *shape)
    u = rnd(g, 1, shape[-1])
    exp = O.spectral_norm(w, u, 1)
    got, u_hat, sigma = ops.spectral_norm(rt, dev(rt, w), dev(rt, u.view(-1)), 1)
    check(got, exp, 1e-4, "spectral norm")


@pytest.mark.parametrize("dt", ["f32", "bf16"])
def test_scale_samples(rt, dt):
    """x[i] *= up[i] * mult per sample; factor 1 leaves the sample untouched, factor 0 zero-fills it WITHOUT reading it (the two
    values the hinge loss produces in the merged discriminator backward), anything else is a plain multiply."""
    g = torch.Generator().manual_seed(31)
    n, per = 7, 4 * 10 * 64
    x = rnd(g, n, 4, 10, 64)
    up = torch.tensor([0.5, 0.0, 0.185, 0.5, 0.0, 1.0, 0.5], dtype=torch.float64)        # * mult 2 -> 1, 0, 0.37, 1, 0, 2, 1
    xd = dev(rt, x, F32 if dt == "f32" else BF16)
    xd[1].fill_(float("nan"))                       # a zero factor must not propagate what was there
    before = xd.clone()
    ops.scale_samples_(rt, xd, dev(rt, up), 2.0)
    exp = before.double().cpu() * (up * 2.0).view(-1, 1, 1, 1)
    exp[1] = 0.0
    exp[4] = 0.0
    assert torch.equal(xd[0], before[0]) and torch.equal(xd[3], before[3]) and torch.equal(xd[6], before[6]), "factor 1: untouched"
    assert float(xd[1].float().abs().max()) == 0.0 and float(xd[4].float().abs().max()) == 0.0, "factor 0: zero-filled"
    check(xd, exp, 1e-6 if dt == "f32" else 8e-3, "scale_samples")
