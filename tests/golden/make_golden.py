#!/usr/bin/env python
"""Generates the golden fixtures in this directory from the fp64 CPU oracle (oracle/sgan_oracle.py).

    python tests/golden/make_golden.py            # rewrites tests/golden/*.npz

PARITY UNPINNED: the reference has no tests / golden vectors and TensorFlow cannot be
installed in this image, so these vectors come from the restatement, not from the TF reference.  They (i) freeze
the oracle against drift (tests/test_golden_cpu.py re-derives every vector and compares at 1e-10) and (ii) give the
GPU box -- which has no /root/reference and where the oracle is only a checker -- fixed vectors to compare the CUDA
path with (tests/test_golden_gpu.py).  Parameters are NOT stored (G+D+R are 59 M floats): they are regenerated from
the oracle's seeded initialisers; per-tensor fingerprints (sum, L2) of the parameters and of the gradients are stored
instead so a change in the initialisers is caught.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import sgan_oracle as O  # noqa: E402

F64 = torch.float64


def ops_vectors():
    g = torch.Generator().manual_seed(20261018)
    out = {}
    rn = lambda *s: torch.randn(*s, generator=g, dtype=F64)
    # Conv2D SAME / VALID (TF padding rule), Conv2DTranspose SAME with strides (2,2), (2,1); 1x1 stride 2
    x = rn(2, 6, 10, 8); w = rn(3, 3, 8, 16) * 0.2; b = rn(16) * 0.1
    out["conv_x"], out["conv_w"], out["conv_b"] = x, w, b
    out["conv_same"] = O.conv2d(x, w, b, "same")
    w2 = rn(2, 2, 8, 16) * 0.2
    out["conv_w2"] = w2
    out["conv_valid"] = O.conv2d(x, w2, b, "valid")
    wt = rn(3, 3, 16, 8) * 0.2
    out["convT_w"] = wt
    out["convT_22"] = O.conv2d_transpose(x, wt, b, (2, 2))
    out["convT_21"] = O.conv2d_transpose(x, wt, b, (2, 1))
    wt1 = rn(1, 1, 16, 8) * 0.2
    out["convT_w1"] = wt1
    out["convT1_22"] = O.conv2d_transpose(x, wt1, b, (2, 2))
    out["convT1_21"] = O.conv2d_transpose(x, wt1, b, (2, 1))
    # filter bank: integer index path is bit exact
    bank = rn(7, 32, 8192) * 0.05
    y = torch.tensor([[0, 6, 3], [5, 5, 1]])
    z0 = rn(2, 32)
    fb = O.filter_bank(z0, y, bank)
    out["fb_y"], out["fb_z0"], out["fb_out"] = y, z0, fb
    out["fb_bank_seed"] = torch.tensor([20261018])
    out["fb_bank_sum"] = bank.sum().reshape(1)
    out["fb_bank"] = bank[:, :, ::64].contiguous()           # thin slice of the bank (fingerprint only)
    out["fb_index_map"] = torch.tensor([[l, k, *O.filter_bank_index_map(l, k)] for l in (0, 1, 2) for k in
                                        (0, 1, 3, 4, 2047, 2048, 4099, 8191)])
    # CTC through K.ctc_batch_cost (eps + re-softmax), blank = C-1, incl. repeated letters and L=1
    probs = torch.softmax(rn(4, 11, 9), -1)
    labels = torch.tensor([[1, 1, 2], [0, 7, 7], [3, 4, 5], [2, 2, 2]])
    il, ll = torch.full((4, 1), 11), torch.full((4, 1), 3)
    out["ctc_probs"], out["ctc_labels"] = probs, labels
    out["ctc_loss"] = O.ctc_batch_cost(labels, probs, il, ll)
    # losses + gradient balancing on (B,1) vectors
    dr, df, sr, sf, s5 = rn(6, 1), rn(6, 1), rn(6, 1), rn(6, 1), rn(6, 1)
    out["loss_in"] = torch.cat([dr, df, sr, sf, s5], 1)
    out["hinge"] = torch.cat(O.hinge(dr, df, sr, sf), 1)
    out["not_saturating"] = torch.cat(O.not_saturating(dr, df, sr, sf, s5), 1)
    r, gl = rn(6, 1).abs() * 10, rn(6, 1)
    gb, rb, _, rs, gs = O.apply_gradient_balancing(r, gl, 1.0)
    out["bal_in"] = torch.cat([r, gl], 1)
    out["bal_out"] = torch.cat([gb, rb], 1)
    out["bal_std"] = torch.stack([rs, gs])
    # spectral norm with explicit u; Conv2D layout (last dim Cout) and Dense
    wsn, u = rn(3, 3, 8, 16), rn(1, 16)
    out["sn_w"], out["sn_u"], out["sn_out"] = wsn, u, O.spectral_norm(wsn, u, 1)
    # Keras Adam / RMSprop, two consecutive steps
    w0, g1, g2 = rn(50), rn(50), rn(50)
    w1, m1, v1 = O.adam_update(w0, g1, torch.zeros(50, dtype=F64), torch.zeros(50, dtype=F64), 1)
    w2_, m2, v2 = O.adam_update(w1, g2, m1, v1, 2)
    out["adam_in"] = torch.stack([w0, g1, g2])
    out["adam_out"] = torch.stack([w1, w2_, m2, v2])
    r1, ms1 = O.rmsprop_update(w0, g1, torch.zeros(50, dtype=F64))
    out["rmsprop_out"] = torch.stack([r1, ms1])
    # BN train: x_hat, new moving stats (Bessel-corrected variance into the moving average)
    xb = rn(3, 4, 5, 6) * 2 + 1
    xh, nm, nv = O.batchnorm_train(xb, torch.zeros(6, dtype=F64), torch.ones(6, dtype=F64))
    out["bn_x"], out["bn_xhat"], out["bn_moving"] = xb, xh, torch.stack([nm, nv])
    return {k: v.numpy() for k, v in out.items()}


def fingerprint(d):
    names = sorted(d)
    return names, np.array([[float(d[k].double().sum()), float(d[k].double().norm())] for k in names])


def train_step_vectors(b=2, l_r=2, l_f=2, loss_fn="hinge", sigma=0.1):
    """One full Mode-A step (G+D+R, gradient balancing on) at a size the fp64 oracle finishes in seconds."""
    g = torch.Generator().manual_seed(7)
    P = {"G": O.make_generator_params(11, F64, sigma=sigma), "D": O.make_discriminator_params(12, F64, sigma=sigma),
         "R": O.make_recognizer_params(13, F64)}
    images = torch.rand(b, 32, 16 * l_r, 1, generator=g, dtype=F64) * 2 - 1
    labels = torch.randint(0, 52, (b, l_r), generator=g)
    fake = torch.randint(0, 52, (b, l_f), generator=g)
    z = torch.randn(b, 128, generator=g, dtype=F64)
    stats, newp, _, grads, extra = O.train_step(P, {}, images, labels, fake, z, loss_fn=loss_fn, apply_gradient_balance=True,
                                                return_grads=True)
    out = {"images": images.numpy(), "labels": labels.numpy(), "fake_labels": fake.numpy(), "z": z.numpy(),
           "seeds": np.array([11, 12, 13]), "sigma": np.array([sigma]),
           "stats": np.array([stats[k] for k in O.STAT_NAMES]),
           "gen_images": extra["gen_images"].numpy(), "d_fake": extra["d_fake"].numpy(), "d_real": extra["d_real"].numpy(),
           "r_fake": extra["r_fake"].numpy(), "r_real": extra["r_real"].numpy()}
    for net in ("G", "D", "R"):
        names, fp = fingerprint(P[net])
        out["param_fp_" + net] = fp
        names_g, fpg = fingerprint(grads[net])
        out["grad_names_" + net] = np.array(names_g)
        out["grad_fp_" + net] = fpg
        _, fpn = fingerprint({k: newp[net][k] for k in names_g})
        out["new_param_fp_" + net] = fpn
    # a few small gradients in full (biases, dense heads, sigma) so that element-wise parity is checked too
    for net, keys in (("D", ("dense.w", "B1.attn.sigma", "B4.conv2.b")), ("R", ("dense.b", "conv7.b")),
                      ("G", ("out.b", "bn.gamma", "B3.attn.sigma", "B1.cbn1.gamma.w"))):
        for k in keys:
            out["grad_{}_{}".format(net, k)] = grads[net][k].numpy()
    return out


def main():
    np.savez_compressed(os.path.join(HERE, "ops_small.npz"), **ops_vectors())
    np.savez_compressed(os.path.join(HERE, "train_step_b2_l2_hinge.npz"), **train_step_vectors())
    np.savez_compressed(os.path.join(HERE, "train_step_b2_l3x1_hinge.npz"), **train_step_vectors(2, 3, 1))
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")


if __name__ == "__main__":
    main()
