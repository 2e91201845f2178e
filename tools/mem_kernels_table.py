#!/usr/bin/env python
"""Achieved HBM bandwidth of the memory-bound kernels of one train step (Mode A, B=64, L=5, bf16; D and R see the fused
[fake;real] batch of 128 images): ALGORITHMIC bytes (each operand read once, each result written once) / ncu launch time
from profiles/r01_launches_step.csv, against the measured copy bandwidth in MEASURED_PEAKS.json.
    python tools/mem_kernels_table.py profiles/r01_launches_step.csv > profiles/r01_memory_bound_kernels.md"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
peak = 6555.8
p = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(p):
    peak = json.load(open(p)).get("hbm_gbs", peak)
lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
L = []
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    L.append((row["Kernel Name"].split("(")[0].replace("void ", ""), float(row["Metric Value"]) / 1e3))


def nth(prefix, which):
    hits = [(i, t) for i, (k, t) in enumerate(L) if k.startswith(prefix)]
    hits.sort(key=lambda x: -x[1])
    return hits[which][1] if len(hits) > which else None


MB = 1e6
rows = []
# (label, kernel prefix, rank by duration among its launches, algorithmic bytes, how counted)
P_D, P_G, P_R = 37336384, 13631488 + 2582530, 5578037
rows.append(("Adam, D (37.3 M params)", "k_adam", 0, P_D * 30, "r w,g,m,v + w w,m,v fp32 + bf16 mirror"))
rows.append(("Adam, G-core + bank (16.2 M)", "k_adam", 1, P_G * 30, "same"))
rows.append(("Adam, R (5.6 M)", "k_adam", 2, P_R * 30, "same"))
px = 128 * 32 * 80
rows.append(("Cin=1 conv D.B1.conv1 fwd (128 img)", "k_conv_fwd_cin1<__nv_bfloat16>", 0, px * 4 + px * 64 * 2, "image fp32 + out bf16"))
rows.append(("Cout=1 conv G.out fwd (64 img)", "k_conv_fwd_cout1", 2, (px // 2) * 64 * 2 + (px // 2) * 4, "act bf16 + out fp32"))
rows.append(("Cin=1 wgrad D.B1.conv1 (128 img)", "k_wgrad_narrow64", 0, px * 64 * 2 + px * 4, "dy bf16 + image fp32"))
rows.append(("max-pool fwd R.p1 (128 img, 64 ch)", "k_maxpool_fwd_v4<__nv_bfloat16>", 0, px * 64 * 2 * 1.25, "in + out bf16"))
rows.append(("avg-pool bwd D.B2 (128 img, 512 ch)", "k_avgpool2_bwd", 0, 128 * 8 * 20 * 512 * 4 + 128 * 16 * 40 * 512 * 2, "dout fp32 + dx bf16"))
rows.append(("BN apply G final (64 img, 32x80x64)", "k_bn_apply", 0, (px // 2) * 64 * (4 + 2), "x fp32 + y bf16"))
rows.append(("BN bwd reduce G final", "k_bn_bwd_reduce<__nv_bfloat16>", 0, (px // 2) * 64 * (4 + 2 + 4), "dy fp32 + act bf16 + x fp32"))
rows.append(("bias-gradient column sum, D.B1 (128 img, 64 ch)", "k_colsum_v4<__nv_bfloat16", 0, px * 64 * 2, "dy bf16"))
rows.append(("filter bank fwd (B=64, L=5)", "k_fb_fwd", 0, 52 * 32 * 8192 * 4 + 64 * 4 * 20 * 512 * 4, "each present character's bank rows once + out"))
rows.append(("filter bank bwd", "k_fb_bwd_bank", 0, 64 * 4 * 20 * 512 * 4 + 52 * 32 * 8192 * 4, "dout + dbank written"))
rows.append(("non-local projections fwd, G (x -> theta,phi,g)", "k_nl_rowgemm<64, 48, 0>", 0, (px // 2) * (64 + 48) * 4, "x + outputs fp32"))
print("| kernel launch | ncu time (us) | algorithmic MB | achieved GB/s | of measured HBM peak (%.0f GB/s) | bytes counted |" % peak)
print("|---|---|---|---|---|---|")
for label, pref, which, nbytes, how in rows:
    t = nth(pref, which)
    if t is None:
        continue
    gbs = nbytes / (t * 1e-6) / 1e9
    print("| %s | %.1f | %.1f | %.0f | %.0f %% | %s |" % (label, t, nbytes / MB, gbs, 100 * gbs / peak, how))
print()
print("Launch times are ncu's (cold L2, serialised, ~2-3 us of fixed cost per launch), so these fractions are lower bounds.")
