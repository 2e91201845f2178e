"""Diagnostic (GPU box): per-tensor gradient errors of the train step vs the fp64 oracle, for each precision mode."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import sgan_oracle as O
runtime = importlib.import_module("scrabble-gan_b200.runtime")
na = importlib.import_module("scrabble-gan_b200.bigacgan.net_architecture")
du = importlib.import_module("scrabble-gan_b200.bigacgan.data_utils")
nl = importlib.import_module("scrabble-gan_b200.bigacgan.net_loss")
optim = importlib.import_module("scrabble-gan_b200.optim")
rt = runtime.Runtime(device=0, mode="fp32"); runtime.set_runtime(rt)
IN_DIM = (32, 160, 1)

def run(mode, use_w, loss_name, balance, b, l_r, l_f, style_encoder, seed=5, brief=False):
    rt.set_mode(mode)
    dt = torch.float64
    g = torch.Generator().manual_seed(seed)
    P = {"G": O.make_generator_params(21, dt, sigma=0.2, bias_scale=0.05, style_encoder_too=style_encoder),
         "D": O.make_discriminator_params(22, dt, sigma=0.2, bias_scale=0.05), "R": O.make_recognizer_params(23, dt, bias_scale=0.05)}
    if use_w: P["W"] = O.make_discriminator_params(24, dt, sigma=0.2, bias_scale=0.05)
    images = torch.rand(b, 32, 16 * l_r, 1, generator=g, dtype=dt) * 2 - 1
    labels = torch.randint(0, 52, (b, l_r), generator=g); fake_labels = torch.randint(0, 52, (b, l_f), generator=g)
    z = torch.randn(b, 128, generator=g, dtype=dt); style = torch.rand(b, 32, 160, 1, generator=g, dtype=dt) * 2 - 1
    g_in = style if style_encoder else z
    stats, newp, newo, grads, extra = O.train_step(P, {}, images, labels, fake_labels, g_in, loss_fn=loss_name, apply_gradient_balance=balance,
                                                   use_style_encoder=style_encoder, use_style_promoter=use_w, return_grads=True, style_images=style)
    # fp32 oracle for a noise floor
    P32 = {n: {k: v.float() for k, v in d.items()} for n, d in P.items()}
    _, _, _, grads32, _ = O.train_step(P32, {}, images.float(), labels, fake_labels, g_in.float(), loss_fn=loss_name, apply_gradient_balance=balance,
                                       use_style_encoder=style_encoder, use_style_promoter=use_w, return_grads=True, style_images=style.float())
    G = na.make_generator(128, IN_DIM, (32, 8192), None, "B3", 52, vis_model=False, rt=rt, style_encoder=style_encoder)
    D = na.make_discriminator(IN_DIM, None, "B1", vis_model=False, rt=rt); R = na.make_recognizer(IN_DIM, None, 53, vis_model=False, rt=rt)
    W = na.make_style_promoter(IN_DIM, None, "B1", vis_model=False, rt=rt) if use_w else None
    G.load_state_dict(P["G"]); D.load_state_dict(P["D"]); R.load_state_dict(P["R"])
    if use_w: W.load_state_dict(P["W"])
    gan = na.make_gan(G, D, R, W, vis_model=False)
    g_opt, d_opt, r_opt, w_opt, loss_fn, disc_iters, agb = optim.setup_optimizer(2e-4, 2e-4, 2e-4, 2e-4, 0.0, 0.999, getattr(nl, loss_name), 1, int(balance), 0)
    out = du.train_step(0, 0, 1, images.float().numpy(), labels.numpy(), D, R, W, gan, g_opt, d_opt, r_opt, w_opt,
                        [s.numpy() for s in style.float()] if (style_encoder or use_w) else None, b, 128, loss_fn, disc_iters, agb,
                        None, 10, "", fake_labels=fake_labels.numpy(), noise=None if style_encoder else z.float().numpy())
    got = dict(zip(du.STAT_NAMES, out))
    print("=== mode", mode, "use_w", use_w, loss_name, "style_enc", style_encoder, "seed", seed)
    if not brief:
        for k in O.STAT_NAMES:
            print("  stat %-18s got %+.6e exp %+.6e" % (k, got[k], stats[k]))
    models = {"G": G, "D": D, "R": R}
    if use_w: models["W"] = W
    for n, m in models.items():
        gd = m.store.grad_dict()
        rows = []
        for k, e in grads[n].items():
            a = gd[k].double().cpu(); e = e.double(); e32 = grads32[n][k].double()
            mx = float(e.abs().max()) + 1e-30
            rows.append((float((a - e).abs().max()) / mx, float((a - e).norm() / (e.norm() + 1e-30)), float((e32 - e).abs().max()) / mx, mx, k))
        rows = [r for r in rows if not r[4].endswith(".up.b")]
        rows.sort(reverse=True)
        if brief:
            print("  net %s worst maxrel %.2e (%s) worst l2rel %.2e" % (n, rows[0][0], rows[0][4], max(r[1] for r in rows)))
            continue
        print(" net", n, "worst tensors (maxrel, l2rel, fp32-oracle maxrel, max|exp|):")
        for r in rows[:14]:
            print("   %.3e %.3e %.3e %.3e %s" % r)

if __name__ == "__main__":
    which = sys.argv[1:] or ["ns", "fork", "tf32", "bf16"]
    if "ns" in which: run("fp32", True, "not_saturating", False, 2, 2, 2, False)
    if "fork" in which: run("fp32", True, "hinge", True, 2, 2, 2, True)
    if "tf32" in which: run("tf32", False, "hinge", True, 3, 2, 3, False)
    if "bf16" in which: run("bf16", False, "hinge", True, 3, 2, 3, False)
    if "matrix" in which:
        run("fp32", True, "hinge", True, 2, 2, 2, False, seed=8)      # W, no style encoder
        run("fp32", False, "hinge", True, 2, 2, 2, True, seed=8)      # style encoder, no W
        run("fp32", True, "hinge", False, 2, 2, 2, False, seed=8)     # W, no balancing
    if "seeds" in which:
        for sd in (6, 7, 8, 9):
            run("fp32", True, "not_saturating", False, 2, 2, 2, False, seed=sd, brief=True)
            run("fp32", True, "hinge", True, 2, 2, 2, True, seed=sd, brief=True)
    if "bf16big" in which: run("bf16", False, "hinge", True, 16, 5, 5, False)
