#!/usr/bin/env python
"""Per-layer timing of the tensor-core conv kernels (fwd / dgrad / wgrad) on the shapes of the D/R/G stacks.
CUDA events on the launching stream, warm-up, L2 flushed between repetitions by rotating over distinct buffers.
    python tools/bench_conv.py [--batch 128] [--reps 20] [--only NAME]
Prints one line per (layer, role): time, algorithmic TFLOP/s, fraction of the measured sustained bf16 peak."""
import argparse
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

ops = importlib.import_module("scrabble-gan_b200.ops")
runtime = importlib.import_module("scrabble-gan_b200.runtime")
abi = importlib.import_module("scrabble-gan_b200._abi")
BF16, F32 = abi.SG_BF16, abi.SG_F32

# name, h, w, ci, co, k  (per image; batch is multiplied in)  -- D at L=5 on the fused [fake;real] batch
LAYERS = [
    ("D.B1.conv2", 32, 80, 64, 64, 3),
    ("D.B2.conv1", 16, 40, 64, 512, 3),
    ("D.B2.conv2", 16, 40, 512, 512, 3),
    ("D.B2.short", 16, 40, 64, 512, 1),
    ("D.B3.conv1", 8, 20, 512, 1024, 3),
    ("D.B3.conv2", 8, 20, 1024, 1024, 3),
    ("D.B3.short", 8, 20, 512, 1024, 1),
    ("D.B4.conv1", 4, 10, 1024, 1024, 3),
    ("D.B4.short", 4, 10, 1024, 1024, 1),
    ("R.conv2", 16, 40, 64, 128, 3),
    ("R.conv4", 8, 20, 256, 256, 3),
    ("R.conv6", 4, 20, 512, 512, 3),
    ("G.B1.conv", 8, 40, 256, 256, 3),
    ("G.B3.conv", 32, 80, 64, 64, 3),
    ("R.conv3", 8, 40, 128, 256, 3),
    ("R.conv5", 4, 40, 256, 512, 3),
    ("G.B2.conv", 16, 80, 128, 128, 3),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--only", default="")
    ap.add_argument("--roles", default="fwd,dgrad,wgrad")
    args = ap.parse_args()
    rt = runtime.Runtime(device=0, mode="bf16")
    runtime.set_runtime(rt)
    peak = 1375.4
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = json.load(open(p)).get("bf16_tflops_sustained", peak)
    n = args.batch
    nbuf = 3
    for name, h, w, ci, co, k in LAYERS:
        if args.only and args.only not in name:
            continue
        flops = 2.0 * n * h * w * k * k * ci * co
        xs = [torch.randn(n, h, w, ci, device=rt.device).to(torch.bfloat16) for _ in range(nbuf)]
        dys = [torch.randn(n, h, w, co, device=rt.device).to(torch.bfloat16) for _ in range(nbuf)]
        wm = torch.randn(k, k, ci, co, device=rt.device) * 0.05
        dw = torch.zeros_like(wm)
        out_f = rt.empty((n, h, w, co), BF16)
        out_d = rt.empty((n, h, w, ci), BF16)
        df = ops.desc_conv_fwd(n, h, w, ci, co, k, k, "same", BF16, BF16, relu=1)
        dd = ops.desc_conv_dgrad(n, h, w, ci, co, k, k, "same", BF16, BF16)
        dwg = ops.desc_conv_fwd(n, h, w, ci, co, k, k, "same", BF16, BF16)
        dacc = ops.desc_conv_fwd(n, h, w, ci, co, k, k, "same", BF16, F32, accumulate=1)      # shortcut: += into the fp32 stream
        dmask = ops.desc_conv_dgrad(n, h, w, ci, co, k, k, "same", BF16, BF16, mask_dt=BF16)  # dgrad gated by the ReLU mask
        out_acc = torch.zeros(n, h, w, co, device=rt.device)
        wf, wd = ops.pack_weights(rt, df, wm), ops.pack_weights(rt, dd, wm)
        bias = torch.zeros(co, device=rt.device)

        def run(role, i):
            if role == "fwd":
                ops.conv_run(rt, df, xs[i % nbuf], wm, wf, bias, None, out_f)
            elif role == "dgrad":
                ops.conv_run(rt, dd, dys[i % nbuf], wm, wd, None, None, out_d)
            elif role == "acc":
                ops.conv_run(rt, dacc, xs[i % nbuf], wm, wf, None, None, out_acc)
            elif role == "dmask":
                ops.conv_run(rt, dmask, dys[i % nbuf], wm, wd, None, xs[(i + 1) % nbuf], out_d)
            else:
                ops.conv_wgrad(rt, dwg, xs[i % nbuf], dys[i % nbuf], dw)
        for role in args.roles.split(","):
            for i in range(3):
                run(role, i)
            torch.cuda.synchronize()
            evs = []
            for i in range(args.reps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                run(role, i)
                e1.record()
                evs.append((e0, e1))
            torch.cuda.synchronize()
            ts = sorted(a.elapsed_time(b) for a, b in evs)
            med = ts[len(ts) // 2]
            tf = flops / (med * 1e-3) / 1e12
            print("%-11s %-5s M=%6d K=%5d N=%4d  %8.1f us  %7.1f TF/s  %5.1f%% of %.0f" % (
                name, role, n * h * w, k * k * ci, co, med * 1e3, tf, 100 * tf / peak, peak), flush=True)


if __name__ == "__main__":
    main()
