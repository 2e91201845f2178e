"""Ragged-width batching: words of different lengths inside ONE launch.
  * CTC with per-sample input_length / label_length -- K.ctc_batch_cost's own interface, which the reference model takes as
    inputs 3 and 4 (net_architecture.py:57-75) but only ever feeds with constants (data_utils.py:400-415);
  * the recogniser on a padded rectangular batch with per-sample lengths (exactly what the reference model computes when it
    is given such a batch)."""
import importlib

import pytest
import torch

import sgan_oracle as O
from _parity import IN_DIM, assert_grads, na, rel_elementwise, rel_max

pytestmark = pytest.mark.gpu
ops = importlib.import_module("scrabble-gan_b200.ops")


def _ragged_case(seed, b, l_max, c):
    g = torch.Generator().manual_seed(seed)
    lens = torch.randint(1, l_max + 1, (b,), generator=g)
    lens[0], lens[1] = l_max, 1                                # both extremes present
    labels = torch.randint(0, c - 1, (b, l_max), generator=g)
    for i in range(b):
        labels[i, lens[i]:] = 0                                # padding value is arbitrary: it must not be read
    labels[2, :lens[2]] = labels[2, 0]                         # a word of one repeated letter (needs 2L - 1 frames)
    return lens, labels, g


@pytest.mark.parametrize("b,l_max,c", [(9, 5, 53), (64, 10, 81)])
def test_ctc_ragged(rt, b, l_max, c):
    lens, labels, g = _ragged_case(5, b, l_max, c)
    t_max = 4 * l_max - 1
    in_len = 4 * lens - 1
    z = torch.randn(b, t_max, c, generator=g, dtype=torch.float64).requires_grad_(True)
    loss = O.ctc_batch_cost(labels, torch.softmax(z, dim=-1), in_len.view(-1, 1), lens.view(-1, 1))
    loss.sum().backward()
    got, grad = ops.ctc(rt, z.detach().float().to(rt.device), labels.to(rt.device, torch.int32), True,
                        in_len.to(rt.device, torch.int32), lens.to(rt.device, torch.int32))
    assert rel_elementwise(got, loss.view(-1), 1.0) <= 1e-4, "ragged CTC loss beyond 1e-4 relative"
    assert rel_max(grad, z.grad) <= 1e-4
    for i in range(b):
        assert float(grad[i, int(in_len[i]):].abs().max() if in_len[i] < t_max else 0.0) == 0.0, "frames beyond T_b must get no gradient"
    # each sample of the ragged batch == the same sample run alone at its own size (the per-bucket result)
    for i in (0, 1, 2, b - 1):
        ti, li = int(in_len[i]), int(lens[i])
        one, gone = ops.ctc(rt, z.detach()[i:i + 1, :ti].float().contiguous().to(rt.device), labels[i:i + 1, :li].contiguous().to(rt.device, torch.int32))
        assert torch.equal(one, got[i:i + 1]) and torch.equal(gone[0], grad[i, :ti]), "ragged and per-bucket CTC must agree bit for bit"


def test_recognizer_on_ragged_batch(rt):
    rt.set_mode("fp32")
    b, l_max = 6, 4
    lens, labels, g = _ragged_case(8, b, l_max, 53)
    P = O.make_recognizer_params(31, torch.float64, bias_scale=0.1)
    x = torch.rand(b, 32, 16 * l_max, 1, generator=g, dtype=torch.float64) * 2 - 1
    for i in range(b):
        x[i, :, 16 * int(lens[i]):] = 1.0                       # white padding right of the word, as a padded IAM batch has
    leaf = {k: v.clone().requires_grad_(not k.endswith(O.NON_TRAINABLE_SUFFIXES)) for k, v in P.items()}
    loss = O.recognizer(x, labels, (4 * lens - 1).view(-1, 1), lens.view(-1, 1), leaf)
    loss.sum().backward()
    R = na.make_recognizer(IN_DIM, None, 53, vis_model=False, rt=rt, initialise=False)
    R.load_state_dict(P)
    # through the model's call signature [images, labels, input_length, label_length] (reference net_architecture.py:66-75)
    out = R([x.float().numpy(), labels.numpy(), (4 * lens - 1).view(-1, 1).numpy(), lens.view(-1, 1).numpy()])
    assert rel_elementwise(out.view(-1), loss.view(-1), 1.0) <= 1e-4
    R.store.zero_grad()
    il, ll = (4 * lens - 1).to(rt.device, torch.int32), lens.to(rt.device, torch.int32)
    got, cache = R.forward(rt, x.float().to(rt.device), labels.to(rt.device, torch.int32), True, il, ll)
    R.backward(rt, cache, None, wgrad=True, want_dx=False)
    assert_grads(R.store.grad_dict(), {k: v.grad for k, v in leaf.items() if v.grad is not None}, 1e-3, 1e-2, "R gradients on a ragged batch")


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_generator_on_ragged_batch(rt, mode):
    """Words of different lengths in ONE generator launch (labels padded with -1) == each word generated alone at its own
    width: the reference's inference loop (run_inference.py:35-50 calls the generator once per word) as one padded batch.
    Also against the fp64 oracle for the words alone."""
    rt.set_mode(mode)
    try:
        b, l_max = 7, 6
        g = torch.Generator().manual_seed(21)
        lens = torch.tensor([6, 1, 3, 2, 6, 4, 5])
        labels = torch.randint(0, 52, (b, l_max), generator=g)
        for i in range(b):
            labels[i, lens[i]:] = -1
        z = torch.randn(b, 128, generator=g, dtype=torch.float64)
        P = O.make_generator_params(40, torch.float64, sigma=0.2, bias_scale=0.05)
        for k in P:                                             # inference BN: moving statistics that matter
            if k.endswith("moving_mean"):
                P[k] = torch.randn(P[k].shape, generator=g, dtype=torch.float64) * 0.1
            if k.endswith("moving_var"):
                P[k] = 0.5 + torch.rand(P[k].shape, generator=g, dtype=torch.float64)
        G = na.make_generator(128, IN_DIM, (32, 8192), None, "B3", 52, vis_model=False, rt=rt, initialise=False)
        G.load_state_dict(P)
        out = G([z.float().numpy(), labels.numpy()], training=False)          # -1 in a host label matrix: ragged
        assert tuple(out.shape) == (b, 32, 16 * l_max, 1)
        n0 = rt.launch_count()
        out_dev = G([z.float().to(rt.device), labels.to(rt.device, torch.int32)], training=False, ragged=True)
        launches = rt.launch_count() - n0
        assert torch.equal(out, out_dev)
        tol = 1e-4 if mode == "fp32" else 3e-2
        alone_launches = 0
        for i in range(b):
            li = int(lens[i])
            n0 = rt.launch_count()
            alone = G([z[i:i + 1].float().numpy(), labels[i:i + 1, :li].numpy()], training=False)
            alone_launches += rt.launch_count() - n0
            err = float((out[i, :, :16 * li] - alone[0]).abs().max())
            assert err <= tol, "word {} (length {}): ragged batch vs alone differ by {:.3e}".format(i, li, err)
            assert float(out[i, :, 16 * li:].abs().max() if li < l_max else 0.0) == 0.0, "the image right of the word must be zero"
            if mode == "fp32":
                exp = O.generator(z[i:i + 1], labels[i:i + 1, :li], P, training=False)
                assert rel_max(out[i, :, :16 * li], exp[0]) <= 1e-3
        assert launches < alone_launches / 3, (launches, alone_launches)
    finally:
        rt.set_mode("fp32")
