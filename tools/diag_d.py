"""Diagnostic: discriminator forward/backward parity vs the fp64 oracle for several input widths (fp32 mode)."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import sgan_oracle as O
runtime = importlib.import_module("scrabble-gan_b200.runtime")
na = importlib.import_module("scrabble-gan_b200.bigacgan.net_architecture")
rt = runtime.Runtime(device=0, mode=sys.argv[1] if len(sys.argv) > 1 else "fp32"); runtime.set_runtime(rt)
P = O.make_discriminator_params(11, torch.float64, sigma=0.3, bias_scale=0.1)
D = na.make_discriminator((32, 160, 1), None, "B1", vis_model=False, rt=rt)
D.load_state_dict(P)
for b, w in ((2, 32), (3, 48), (2, 80), (2, 160), (3, 160), (1, 160), (2, 128), (2, 96)):
    g = torch.Generator().manual_seed(0)
    x = (torch.rand(b, 32, w, 1, generator=g, dtype=torch.float64) * 2 - 1).requires_grad_(True)
    up = torch.randn(b, generator=g, dtype=torch.float64)
    leaf = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    logits = O.discriminator(x, leaf, "B1")
    (logits.view(-1) * up).sum().backward()
    D.store.zero_grad()
    got, cache = D.forward(rt, x.detach().float().to(rt.device))
    dx = D.backward(rt, cache, up.float().to(rt.device), wgrad=True, want_dx=True)
    gd = D.store.grad_dict()
    rows = []
    for k, v in leaf.items():
        e = v.grad; a = gd[k].double().cpu()
        rows.append((float((a - e).abs().max() / (e.abs().max() + 1e-30)), k))
    rows.sort(reverse=True)
    fe = float((got.double().cpu().view(-1) - logits.detach().view(-1)).abs().max() / logits.abs().max())
    de = float((dx.double().cpu() - x.grad).abs().max() / x.grad.abs().max())
    print("b=%d w=%d fwd %.2e dx %.2e worst:" % (b, w, fe, de), ["%.2e %s" % r for r in rows[:4]], flush=True)
