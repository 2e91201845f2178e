#!/usr/bin/env python
"""Per-launch timing of every convolution launch of ONE eager train step, in step order, with its shape (rt.trace):
which layers the tensor-core time goes to, and at what fraction of the measured bf16 peak each runs inside the step
(warm caches, neighbours in flight -- unlike ncu's serialised cold-cache list).
    python tools/trace_step.py [--batch 64] [--length 5] [--dtype bf16] [--reps 5]"""
import argparse
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--length", type=int, default=5)
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    runtime = importlib.import_module("scrabble-gan_b200.runtime")
    na = importlib.import_module("scrabble-gan_b200.bigacgan.net_architecture")
    du = importlib.import_module("scrabble-gan_b200.bigacgan.data_utils")
    nl = importlib.import_module("scrabble-gan_b200.bigacgan.net_loss")
    optim = importlib.import_module("scrabble-gan_b200.optim")
    rt = runtime.Runtime(device=0, mode=args.dtype)
    runtime.set_runtime(rt)
    peak = 1375.4
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = json.load(open(p)).get("bf16_tflops_sustained", peak)
    B, L = args.batch, args.length
    in_dim = (32, 160, 1)
    G = na.make_generator(128, in_dim, (32, 8192), None, "B3", 52, vis_model=False, rt=rt)
    D = na.make_discriminator(in_dim, None, "B1", vis_model=False, rt=rt)
    R = na.make_recognizer(in_dim, None, 53, vis_model=False, rt=rt)
    gan = na.make_gan(G, D, R, None, vis_model=False)
    g_opt, d_opt, r_opt, w_opt, loss_fn, disc_iters, agb = optim.setup_optimizer(2e-4, 2e-4, 2e-4, 2e-4, 0.0, 0.999, nl.hinge, 1, 1, 0)
    rng = np.random.RandomState(1)
    imgs = torch.from_numpy(rng.uniform(-1, 1, size=(B, 32, 16 * L, 1)).astype(np.float32)).to(rt.device)
    labels = torch.from_numpy(rng.randint(0, 52, size=(B, L)).astype(np.int32)).to(rt.device)
    fake = torch.from_numpy(rng.randint(0, 52, size=(B, L)).astype(np.int32)).to(rt.device)
    z = torch.from_numpy(rng.standard_normal(size=(B, 128)).astype(np.float32)).to(rt.device)
    du.GRAPH_ENABLED = False

    def step(i):
        return du.train_step(0, i, 1, imgs, labels, D, R, None, gan, g_opt, d_opt, r_opt, w_opt, None, B, 128, loss_fn, disc_iters,
                             agb, None, 10, "", fake_labels=fake, noise=z, return_device_stats=True)
    for i in range(3):
        step(i)
    torch.cuda.synchronize()
    runs = []
    step_ms = []
    for r in range(args.reps):
        rt.trace = []
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step(3 + r)
        e1.record()
        torch.cuda.synchronize()
        step_ms.append(e0.elapsed_time(e1))
        runs.append([(role, d, a.elapsed_time(b) * 1e3) for role, d, a, b in rt.trace])
        rt.trace = None
    n = len(runs[0])
    assert all(len(r) == n for r in runs)
    rows = []
    for i in range(n):
        role, d, _ = runs[0][i]
        us = float(np.median([r[i][2] for r in runs]))
        gflop = 2.0 * d["m"] * d["k"] * d["co"] / 1e9
        rows.append((i, role, d, us, gflop))
    tot = sum(r[3] for r in rows)
    print("eager step {:.2f} ms (median of {}, events around every conv launch); {} conv launches, {:.1f} us in total, {:.1f} GFLOP".format(
        float(np.median(step_ms)), args.reps, n, tot, sum(r[4] for r in rows)))
    print("{:>3s} {:<9s} {:>7s} {:>6s} {:>6s} {:>5s} {:>9s} {:>8s} {:>7s} {:>6s}  {}".format("#", "role", "M", "K", "N", "taps", "grid_hw", "us", "TF/s", "frac", "notes"))
    for i, role, d, us, gflop in rows:
        print("{:>3d} {:<9s} {:>7d} {:>6d} {:>6d} {:>5d} {:>9s} {:>8.1f} {:>7.1f} {:>6.3f}  stride{} in_stride{}{}{}".format(
            i, role, d["m"], d["k"], d["co"], d["taps"], "{}x{}".format(*d["grid_hw"]), us, gflop / us * 1e3, gflop / us * 1e3 / peak,
            d["stride"], d["in_stride"], " relu" if d["relu"] else "", " acc" if d["acc"] else ""))
    by = {}
    for i, role, d, us, gflop in rows:
        e = by.setdefault(role, [0, 0.0, 0.0])
        e[0] += 1; e[1] += us; e[2] += gflop
    for role, (c, us, gf) in by.items():
        print("{:<9s} {:>3d} launches {:>8.1f} us {:>8.1f} GFLOP -> {:.1f} TFLOP/s ({:.3f} of {:.0f})".format(role, c, us, gf, gf / us * 1e3, gf / us * 1e3 / peak, peak))


if __name__ == "__main__":
    main()
