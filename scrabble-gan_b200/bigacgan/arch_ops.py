"""NonLocalBlock, SpatialEmbedding (filter bank) and spectral_norm -- B200-native counterparts of the reference's
src/bigacgan/arch_ops.py (NonLocalBlock :5-72, SpatialEmbedding :77-95, spectral_norm :98-126)."""
from __future__ import annotations

import torch

from .. import ops
from .._abi import SG_F32
from ..params import ParamStore, init_filter_bank, init_orthogonal, init_zeros
from ..runtime import Runtime, get_runtime


class NonLocalBlock:
    """Self-attention block: theta, phi (C/8) and g (C/2) 1x1 projections; phi, g max-pooled 2x2;
    softmax(theta phi^T) over keys (no scaling) @ g; 1x1 back to C; sigma * o + x with sigma a scalar (init 0).

    Deviation D2: the reference re-creates (re-randomises) the four 1x1 kernels on every call and never
    trains them; this block keeps them as persistent trainable variables (the intended SAGAN block).  The q x kv
    attention map is never materialised (flash-style kernels, logsumexp kept for the backward)."""

    def __init__(self, store: ParamStore, name: str, c: int):
        assert c == 64, "the attention kernels are built for C = 64 (dk = 8, dv = 32)"
        self.c, self.dk, self.dv = c, c // 8, c // 2
        self.theta = store.add(name + ".theta.w", (1, 1, c, self.dk), init_orthogonal)
        self.phi = store.add(name + ".phi.w", (1, 1, c, self.dk), init_orthogonal)
        self.g = store.add(name + ".g.w", (1, 1, c, self.dv), init_orthogonal)
        self.o = store.add(name + ".o.w", (1, 1, self.dv, c), init_orthogonal)
        self.sigma = store.add(name + ".sigma", (1,), init_zeros)

    def forward(self, rt: Runtime, x, kv_cols=None):
        """kv_cols (int32 [n], ragged batches): image n attends to the first kv_cols[n] columns of the pooled key map only."""
        n, h, w, c = x.shape
        theta, phi_f, g_f = ops.nonlocal_proj_fwd(rt, x, self.theta.eff, self.phi.eff, self.g.eff)
        phi_f, g_f = phi_f.view(n, h, w, self.dk), g_f.view(n, h, w, self.dv)
        phi = ops.maxpool_fwd(rt, phi_f, 2, 2)
        g = ops.maxpool_fwd(rt, g_f, 2, 2)
        q, kv = h * w, (h // 2) * (w // 2)
        o, lse = ops.attn_fwd(rt, theta.view(n, q, self.dk), phi.view(n, kv, self.dk), g.view(n, kv, self.dv), kv_w=w // 2, kv_cols=kv_cols)
        og, out = ops.nonlocal_out_fwd(rt, o, self.o.eff, self.sigma.data, x)
        # every cached tensor keeps the image index as its first dimension so that sub-batches can be sliced
        return out.view(n, h, w, c), (x, theta.view(n, q, self.dk), phi_f, phi, g_f, g, o, lse, og.view(n, h, w, c))

    @staticmethod
    def slice_cache(cache, a: int, b: int):
        return tuple(t[a:b] for t in cache)

    def backward(self, rt: Runtime, cache, dout, wgrad: bool = True):
        """Returns dx; `dout` is consumed (the identity-path gradient is accumulated in place)."""
        x, theta, phi_f, phi, g_f, g, o, lse, og = cache
        n, h, w, c = x.shape
        q, kv = h * w, (h // 2) * (w // 2)
        if wgrad:
            ops.dot_into(rt, dout, og, self.sigma.grad, accumulate=1)
        d_o = ops.nonlocal_out_bwd(rt, dout, o, self.o.eff, self.sigma.data, self.o.grad if wgrad else None)
        dtheta, dphi, dg = ops.attn_bwd(rt, theta.view(n, q, self.dk), phi.view(n, kv, self.dk), g.view(n, kv, self.dv),
                                        o.view(n, q, self.dv), lse, d_o.view(n, q, self.dv))
        dphi_f = ops.maxpool_bwd(rt, dphi.view(n, h // 2, w // 2, self.dk), phi_f, 2, 2, False, SG_F32)
        dg_f = ops.maxpool_bwd(rt, dg.view(n, h // 2, w // 2, self.dv), g_f, 2, 2, False, SG_F32)
        gr = (self.theta.grad, self.phi.grad, self.g.grad) if wgrad else (None, None, None)
        return ops.nonlocal_proj_bwd(rt, x, dtheta, dphi_f, dg_f, self.theta.eff, self.phi.eff, self.g.eff, dout, *gr)


class SpatialEmbedding:
    """The per-character filter bank: one learned 32 x 8192 matrix per alphabet symbol (reference arch_ops.py:77-95
    builds only the gather; the multiply-by-z0 and the reshape/reshape/transpose assembly of
    net_architecture.py:260-271 are fused here into one kernel that writes NHWC (B,4,4L,512) directly)."""

    def __init__(self, store: ParamStore, vocab_size: int, filter_dim=(32, 8192)):
        assert tuple(filter_dim) == (32, 8192), "embed_y must be (32, 8192) = (latent_dim/4, 512*4*4)"
        self.vocab_size = vocab_size
        self.kernel = store.add("filter_bank", (vocab_size, filter_dim[0], filter_dim[1]), init_filter_bank)

    def forward(self, rt: Runtime, z, z_stride: int, y):
        return ops.filterbank_fwd(rt, z, z_stride, y, self.kernel.data), (z, z_stride, y)

    def backward(self, rt: Runtime, cache, dout, dz_out=None, dz_ld: int = 0):
        z, z_stride, y = cache
        # the kernel overwrites every element of dbank (deterministic, no atomics): exactly one call per step
        ops.filterbank_bwd(rt, dout, z, z_stride, y, self.kernel.data, self.kernel.grad, dz_out, dz_ld)


def spectral_norm(w, power_iteration: int = 1, u=None):
    """Same call shape as the reference's spectral_norm(w, power_iteration=1) (arch_ops.py:99): returns w / sigma
    after `power_iteration` power steps from a random u ~ N(0,1) (pass `u` for reproducibility).
    Accepts a torch CUDA tensor, a Variable, or any DLPack producer; returns a torch CUDA tensor.

    NOTE: in the reference this function is installed as a Keras kernel_regularizer whose loss is never
    read, so it has no effect on the train step; it is offered here as a standalone operator only."""
    from .._abi import from_dlpack
    rt = get_runtime()
    if hasattr(w, "data") and hasattr(w, "store"):
        w = w.data
    w = from_dlpack(w).to(device=rt.device, dtype=torch.float32).contiguous()
    cols = w.shape[-1]
    if u is None:
        u = torch.randn(cols, device=rt.device, dtype=torch.float32)
    else:
        u = from_dlpack(u).to(device=rt.device, dtype=torch.float32).reshape(-1).contiguous()
    w_norm, _, _ = ops.spectral_norm(rt, w, u, power_iteration)
    return w_norm
