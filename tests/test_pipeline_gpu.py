"""Input pipeline and random numbers of the step's caller side (K21): the Philox kernel that
replaces tf.random.normal (data_utils.py:385), the pinned-memory prefetcher around a `load_prepare_data`-style generator
(data_utils.py:14-84), the device-side synthetic loader."""
import importlib

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ops = importlib.import_module("scrabble-gan_b200.ops")
du = importlib.import_module("scrabble-gan_b200.bigacgan.data_utils")


def _philox_ref(ctr, seed):
    """Philox4x32-10 (Salmon et al. 2011) in Python integers: counter (ctr, 0), key = seed."""
    c = [ctr & 0xFFFFFFFF, (ctr >> 32) & 0xFFFFFFFF, 0, 0]
    k = [seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF]
    for _ in range(10):
        p0, p1 = 0xD2511F53 * c[0], 0xCD9E8D57 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k[0]) & 0xFFFFFFFF, p1 & 0xFFFFFFFF, ((p0 >> 32) ^ c[3] ^ k[1]) & 0xFFFFFFFF, p0 & 0xFFFFFFFF]
        k = [(k[0] + 0x9E3779B9) & 0xFFFFFFFF, (k[1] + 0xBB67AE85) & 0xFFFFFFFF]
    return c


def test_philox_known_answer_and_stream(rt):
    # the Random123 known-answer vector for philox4x32-10: counter = key = 0 -> 6627e8d5 e169c58d bc57ac4c 9b00dbd8
    assert _philox_ref(0, 0) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    rt.manual_seed(99)
    u = ops.random_(rt, rt.empty((4096,)), normal=False)
    exp = np.array([x for g in range(1024) for x in _philox_ref(g, 99)], dtype=np.float64) * 4.656612873077393e-10 - 1.0
    assert np.allclose(u.cpu().numpy(), exp.astype(np.float32), atol=1e-6), "uniform draws must be the Philox stream bit for bit"
    assert int(rt.rng_state().item()) == 1024, "the device-side stream position advances by ceil(n / 4)"
    v = ops.random_(rt, rt.empty((4096,)), normal=False)
    exp2 = np.array([x for g in range(1024, 2048) for x in _philox_ref(g, 99)], dtype=np.float64) * 4.656612873077393e-10 - 1.0
    assert np.allclose(v.cpu().numpy(), exp2.astype(np.float32), atol=1e-6)
    rt.manual_seed(99)
    assert torch.equal(u, ops.random_(rt, rt.empty((4096,)), normal=False)), "same seed, same position: same numbers"
    z = ops.random_(rt, rt.empty((1 << 20,))).double()
    assert abs(float(z.mean())) < 5e-3 and abs(float(z.var()) - 1.0) < 5e-3 and abs(float((z ** 3).mean())) < 2e-2
    assert abs(float((z ** 4).mean()) - 3.0) < 5e-2 and float(z.abs().max()) < 7.0
    # captured in a CUDA graph the launch continues the stream on every replay
    out = rt.empty((1024,))
    ops.random_(rt, out)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream(device=rt.device)
    with torch.cuda.graph(g, stream=side):
        rt.use_current_stream()
        ops.random_(rt, out)
    rt.use_current_stream()
    g.replay()
    a = out.clone()
    g.replay()
    assert not torch.equal(a, out)


def test_prefetcher_and_device_loader(rt):
    char_vec = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ"
    ref = du.synthetic_word_batches((32, 160, 1), 8, char_vec, 4, seed=3)
    expect = [next(ref) for _ in range(12)]
    pf = du.DevicePrefetcher(du.synthetic_word_batches((32, 160, 1), 8, char_vec, 4, seed=3), rt, depth=2)
    for imgs, labels in expect:
        d_img, d_lab = next(pf)
        assert d_img.is_cuda and d_lab.is_cuda and d_lab.dtype == torch.int32
        assert np.array_equal(d_img.cpu().numpy(), imgs) and np.array_equal(d_lab.cpu().numpy(), labels)
    pf.close()
    # a finite dataset ends the iteration; an exception in the source surfaces in the consumer
    pf = du.DevicePrefetcher(iter(expect[:3]), rt)
    assert len(list(pf)) == 3

    def broken():
        yield expect[0]
        raise RuntimeError("boom")
    pf = du.DevicePrefetcher(broken(), rt)
    next(pf)
    with pytest.raises(RuntimeError):
        next(pf)
    dev = du.synthetic_word_batches_device(rt, (32, 160, 1), 8, char_vec, 10, seed=1)
    for _ in range(5):
        imgs, labels = next(dev)
        l = labels.shape[1]
        assert imgs.shape == (8, 32, 16 * l, 1) and imgs.is_cuda and float(imgs.min()) >= -1.0 and float(imgs.max()) < 1.0
        assert labels.dtype == torch.int32 and int(labels.min()) >= 0 and int(labels.max()) < 52
