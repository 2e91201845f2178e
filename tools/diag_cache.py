"""Diagnostic: detect corruption of forward caches between forward and backward inside train_step."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
du = importlib.import_module("scrabble-gan_b200.bigacgan.data_utils")
na = importlib.import_module("scrabble-gan_b200.bigacgan.net_architecture")

def flat(c, out, path=""):
    if isinstance(c, torch.Tensor):
        out.append((path, c))
    elif isinstance(c, (tuple, list)):
        for i, e in enumerate(c):
            flat(e, out, path + "/" + str(i))

snap = {}
orig_fwd = na.Discriminator.forward
orig_bwd = na.Discriminator.backward
def fwd(self, rt, x):
    y, cache = orig_fwd(self, rt, x)
    ts = []; flat(cache, ts)
    snap[id(cache)] = [(p, t, t.double().sum().item(), t.double().abs().sum().item()) for p, t in ts]
    return y, cache
def bwd(self, rt, cache, up, wgrad=True, want_dx=False):
    for p, t, s, a in snap[id(cache)]:
        s2, a2 = t.double().sum().item(), t.double().abs().sum().item()
        if s2 != s or a2 != a:
            print("CACHE CHANGED", self.name, p, tuple(t.shape), t.dtype, s, s2, a, a2, flush=True)
    print("bwd", self.name, "up", up.cpu().tolist(), "wgrad", wgrad, "want_dx", want_dx, flush=True)
    return orig_bwd(self, rt, cache, up, wgrad, want_dx)
na.Discriminator.forward = fwd
na.Discriminator.backward = bwd
sys.argv = [sys.argv[0], "ns"]
exec(open(os.path.join(ROOT, "tools", "diag_parity.py")).read())
