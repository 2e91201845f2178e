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


# ---- a DLPack producer that is NOT torch: device memory from the CUDA runtime (cuda-python), capsule built with ctypes --------
class _RawDeviceArray:
    """Minimal DLPack (v0.x `dltensor` capsule) producer over cudaMalloc'ed memory: what a TF2 tensor
    (tf.experimental.dlpack.to_dlpack), a cupy array or any other framework hands over."""
    _keep = {}

    def __init__(self, host: np.ndarray):
        import ctypes as C
        from cuda.bindings import runtime as cudart
        self.C, self.cudart = C, cudart
        host = np.ascontiguousarray(host)
        self.shape, self.dtype, self.nbytes = host.shape, host.dtype, host.nbytes
        err, self.dptr = cudart.cudaMalloc(max(self.nbytes, 1))
        assert int(err) == 0
        (err,) = cudart.cudaMemcpy(self.dptr, host.ctypes.data, self.nbytes, cudart.cudaMemcpyKind.cudaMemcpyHostToDevice)
        assert int(err) == 0

    def __dlpack_device__(self):
        return (2, 0)                                    # kDLCUDA, device 0

    def __dlpack__(self, stream=None, **kwargs):
        C = self.C

        class DLDevice(C.Structure):
            _fields_ = [("device_type", C.c_int32), ("device_id", C.c_int32)]

        class DLDataType(C.Structure):
            _fields_ = [("code", C.c_uint8), ("bits", C.c_uint8), ("lanes", C.c_uint16)]

        class DLTensor(C.Structure):
            _fields_ = [("data", C.c_void_p), ("device", DLDevice), ("ndim", C.c_int32), ("dtype", DLDataType),
                        ("shape", C.POINTER(C.c_int64)), ("strides", C.POINTER(C.c_int64)), ("byte_offset", C.c_uint64)]

        class DLManagedTensor(C.Structure):
            pass
        deleter_t = C.CFUNCTYPE(None, C.POINTER(DLManagedTensor))
        DLManagedTensor._fields_ = [("dl_tensor", DLTensor), ("manager_ctx", C.c_void_p), ("deleter", deleter_t)]
        shape = (C.c_int64 * len(self.shape))(*self.shape)
        mt = DLManagedTensor()
        code, bits = (2, 32) if self.dtype == np.float32 else (0, 32)        # kDLFloat / kDLInt
        mt.dl_tensor = DLTensor(C.c_void_p(int(self.dptr)), DLDevice(2, 0), len(self.shape), DLDataType(code, bits, 1), shape, None, 0)
        key = id(mt)

        def _deleter(_p, key=key):
            _RawDeviceArray._keep.pop(key, None)
        mt.deleter = deleter_t(_deleter)
        _RawDeviceArray._keep[key] = (mt, shape, mt.deleter, self)          # owner and C structs outlive the consumer's view
        C.pythonapi.PyCapsule_New.restype = C.py_object
        C.pythonapi.PyCapsule_New.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
        return C.pythonapi.PyCapsule_New(C.addressof(mt), b"dltensor", None)


def test_non_torch_dlpack_producer_feeds_the_api(rt):
    """The drop-in boundary takes tensors of OTHER frameworks: a generator call and a train step fed with DLPack
    producers that are not torch tensors give exactly what the same data as numpy / torch gives."""
    pytest.importorskip("cuda.bindings.runtime")
    na = importlib.import_module("scrabble-gan_b200.bigacgan.net_architecture")
    nl = importlib.import_module("scrabble-gan_b200.bigacgan.net_loss")
    optim = importlib.import_module("scrabble-gan_b200.optim")
    rt.set_mode("fp32")
    rng = np.random.RandomState(3)
    b, l = 4, 3
    z = rng.standard_normal((b, 128)).astype(np.float32)
    labels = rng.randint(0, 52, (b, l)).astype(np.int32)
    imgs = rng.uniform(-1, 1, (b, 32, 16 * l, 1)).astype(np.float32)

    def build():
        G = na.make_generator(128, (32, 160, 1), (32, 8192), None, "B3", 52, vis_model=False, rt=rt, seed=5)
        D = na.make_discriminator((32, 160, 1), None, "B1", vis_model=False, rt=rt, seed=6)
        R = na.make_recognizer((32, 160, 1), None, 53, vis_model=False, rt=rt, seed=7)
        return G, D, R
    G, D, R = build()
    ref_img = G([z, labels], training=False)
    got_img = G([_RawDeviceArray(z), _RawDeviceArray(labels)], training=False)
    assert torch.equal(ref_img, got_img)

    def one_step(nets, images, lab, fake, noise):
        G, D, R = nets
        gan = na.make_gan(G, D, R, None, vis_model=False)
        g_opt, d_opt, r_opt, w_opt, loss_fn, disc_iters, agb = optim.setup_optimizer(2e-4, 2e-4, 2e-4, 2e-4, 0.0, 0.999, nl.hinge, 1, 1, 0)
        old = du.GRAPH_ENABLED
        du.GRAPH_ENABLED = False
        try:
            return du.train_step(0, 0, 1, images, lab, D, R, None, gan, g_opt, d_opt, r_opt, w_opt, None, b, 128, loss_fn, disc_iters, agb,
                                 None, 10, "", fake_labels=fake, noise=noise)
        finally:
            du.GRAPH_ENABLED = old
    fake = rng.randint(0, 52, (b, l)).astype(np.int32)
    a = one_step(build(), imgs, labels, fake, z)
    c = one_step(build(), _RawDeviceArray(imgs), _RawDeviceArray(labels), _RawDeviceArray(fake), _RawDeviceArray(z))
    assert a == c, "same data through numpy and through a foreign DLPack producer must give the same 16 statistics"
