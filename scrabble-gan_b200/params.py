"""Flat parameter storage.

Each network keeps all trainable parameters in ONE contiguous fp32 buffer (and one gradient buffer of the same
layout) so that (i) the optimizer is a single fused launch per network, (ii) the data-parallel gradient exchange is
one NCCL sum-all-reduce per network over the bucket itself (no packing copy), and (iii) zeroing gradients is one
memset.  Variables are views; names follow the oracle's naming so that weights can be exchanged by name in TF
layouts (HWIO convs, (kh,kw,Cout,Cin) transposed convs, (in,out) dense)."""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Sequence

import torch

ALIGN = 64      # elements; keeps every variable 256-byte aligned (TMA / float4 requirements)


class Variable:
    """A named view into a ParamStore (the analogue of a tf.Variable in `model.trainable_variables`)."""

    def __init__(self, store: "ParamStore", name: str, shape: Sequence[int], trainable: bool, init):
        self.store = store
        self.name = name
        self.shape = tuple(int(s) for s in shape)
        self.trainable = trainable
        self.init = init
        self.offset = -1
        self.numel = 1
        for s in self.shape:
            self.numel *= s

    @property
    def data(self) -> torch.Tensor:
        buf = self.store.w if self.trainable else self.store.s
        return buf[self.offset:self.offset + self.numel].view(self.shape)

    @property
    def grad(self) -> torch.Tensor:
        assert self.trainable
        return self.store.g[self.offset:self.offset + self.numel].view(self.shape)

    @property
    def eff(self) -> torch.Tensor:
        """The values the COMPUTE path uses: the variable itself, or -- with the spectral-norm re-parameterisation switched
        on (ParamStore.enable_spectral_norm) -- its normalised copy W / sigma in the store's effective-weight buffer."""
        buf = self.store.w_eff
        if buf is None or not self.trainable:
            return self.data
        return buf[self.offset:self.offset + self.numel].view(self.shape)

    def mirror(self, rt) -> torch.Tensor:
        """bf16 mirror of this variable (view into the store's mirror buffer, refreshed lazily)."""
        assert self.trainable
        return self.store.mirror(rt)[self.offset:self.offset + self.numel].view(self.shape)

    def numpy(self):
        return self.data.detach().cpu().numpy()

    def assign(self, value) -> None:
        t = torch.as_tensor(value).to(device=self.data.device, dtype=torch.float32).reshape(self.shape)
        self.data.copy_(t)
        self.store.version += 1

    def __repr__(self):
        return "Variable({}, shape={}, trainable={})".format(self.name, self.shape, self.trainable)


# ---- initialisers (construction time only: they run on the HOST -- torch CPU as RNG / LAPACK QR -- and the result is
# copied into the flat buffer, so model construction launches no GPU kernels besides the copies) -------
def init_orthogonal(shape, device, gen):
    """tf.initializers.orthogonal(gain=1): QR of a normal matrix of shape (prod(shape[:-1]), shape[-1])."""
    rows = 1
    for s in shape[:-1]:
        rows *= s
    cols = shape[-1]
    a = torch.randn(max(rows, cols), min(rows, cols), generator=gen, device=device, dtype=torch.float32)
    q, r = torch.linalg.qr(a)
    q = q * torch.sign(torch.diagonal(r))
    if rows < cols:
        q = q.t()
    return q.reshape(shape).contiguous()


def init_glorot_uniform(shape, device, gen):
    rf = 1
    for s in shape[:-2]:
        rf *= s
    fan_in, fan_out = shape[-2] * rf, shape[-1] * rf
    lim = math.sqrt(6.0 / (fan_in + fan_out))
    return (torch.rand(shape, generator=gen, device=device, dtype=torch.float32) * 2 - 1) * lim


def init_filter_bank(shape, device, gen):
    """Keras add_weight default (glorot_uniform) on a rank-3 [vocab, 32, 8192] weight: receptive field = vocab."""
    fan_in, fan_out = shape[-2] * shape[0], shape[-1] * shape[0]
    lim = math.sqrt(6.0 / (fan_in + fan_out))
    return (torch.rand(shape, generator=gen, device=device, dtype=torch.float32) * 2 - 1) * lim


def init_zeros(shape, device, gen):
    return torch.zeros(shape, device=device, dtype=torch.float32)


def init_ones(shape, device, gen):
    return torch.ones(shape, device=device, dtype=torch.float32)


class ParamStore:
    def __init__(self, rt, name: str, seed: int = 0):
        self.rt = rt
        self.name = name
        self.seed = seed
        self.vars: List[Variable] = []
        self.by_name: Dict[str, Variable] = {}
        self.w: Optional[torch.Tensor] = None      # trainable values   (flat fp32)
        self.g: Optional[torch.Tensor] = None      # gradients          (flat fp32)
        self.s: Optional[torch.Tensor] = None      # non-trainable state (BN moving statistics)
        self.version = 0                           # bumped whenever values change (packed-weight caches key on it)
        self.wb: Optional[torch.Tensor] = None     # bf16 mirror of w (same offsets), kept current by the optimizer kernel
        self.wb_version = -1
        self.n_trainable = 0
        self.w_eff: Optional[torch.Tensor] = None  # effective weights (spectral-norm re-parameterisation on): same layout as w
        self.sn: Optional["SpectralNormState"] = None

    def add(self, name: str, shape, init: Callable = init_zeros, trainable: bool = True) -> Variable:
        assert self.w is None, "store already finalised"
        assert name not in self.by_name, "duplicate variable " + name
        v = Variable(self, name, shape, trainable, init)
        self.vars.append(v)
        self.by_name[name] = v
        return v

    def finalize(self, initialise: bool = True) -> None:
        off_w = off_s = 0
        for v in self.vars:
            if v.trainable:
                v.offset = off_w
                off_w += (v.numel + ALIGN - 1) // ALIGN * ALIGN
            else:
                v.offset = off_s
                off_s += (v.numel + ALIGN - 1) // ALIGN * ALIGN
        dev = self.rt.device
        self.w = torch.zeros(max(off_w, 1), device=dev, dtype=torch.float32)
        self.g = torch.zeros(max(off_w, 1), device=dev, dtype=torch.float32)
        self.s = torch.zeros(max(off_s, 1), device=dev, dtype=torch.float32)
        self.n_trainable = sum(v.numel for v in self.vars if v.trainable)
        if initialise:
            gen = torch.Generator().manual_seed(self.seed)
            for v in self.vars:
                v.data.copy_(v.init(v.shape, "cpu", gen))
        else:
            # weights will be loaded by the caller (load_state_dict / load_weights): only the non-zero constants
            for v in self.vars:
                if v.init is init_ones:
                    v.data.fill_(1.0)
        self.version += 1

    def mirror(self, rt) -> torch.Tensor:
        """The bf16 mirror of the flat weight buffer.  The fused Adam launch writes it together with the fp32 weights;
        after any other change (initialisation, weight load, assign, per-variable updates) it is re-cast here."""
        if self.wb is None:
            self.wb = torch.empty_like(self.w, dtype=torch.bfloat16)
        if self.wb_version != self.version:
            from . import ops
            from ._abi import SG_BF16
            src = self.w if self.w_eff is None else self.w_eff
            ops.call.sg_cast(rt.ctx, ops._p(src), ops._p(self.wb), SG_BF16, self.w.numel())
            self.wb_version = self.version
        return self.wb

    def enable_spectral_norm(self, names, seed: int = 0) -> "SpectralNormState":
        """Switch on the spectral-norm weight re-parameterisation for the listed kernels (paper-faithful option; in the
        reference spectral_norm is a kernel_regularizer nobody reads -- SURVEY Q2).  The compute path then reads
        W / sigma(W) from `w_eff`; one power-iteration step per training forward with a PERSISTENT u."""
        assert self.w is not None, "finalize the store first"
        self.w_eff = self.w.clone()
        self.sn = SpectralNormState(self, [self.by_name[n] for n in names], seed)
        self.version += 1
        return self.sn

    @property
    def trainable_variables(self) -> List[Variable]:
        return [v for v in self.vars if v.trainable]

    def zero_grad(self) -> None:
        from . import ops
        ops.call.sg_zero(self.rt.ctx, ops._p(self.g), self.g.numel() * 4)       # a memset node, not a (torch) fill kernel

    def state_dict(self) -> Dict[str, torch.Tensor]:
        return {v.name: v.data.detach().clone() for v in self.vars}

    def load_state_dict(self, sd, strict: bool = True) -> None:
        for v in self.vars:
            if v.name in sd:
                v.data.copy_(torch.as_tensor(sd[v.name]).to(device=self.w.device, dtype=torch.float32).reshape(v.shape))
            elif strict:
                raise KeyError("missing weight " + v.name)
        self.version += 1

    def grad_dict(self) -> Dict[str, torch.Tensor]:
        return {v.name: v.grad.detach().clone() for v in self.vars if v.trainable}


class SpectralNormState:
    """W_sn = W / sigma, sigma = v^T W u after one power-iteration step from the persistent u (arch_ops.py:99-126 computes the
    same quantity from a fresh random u; SN-GAN / compare_gan keep u).  forward() refreshes the effective weights of the
    store, backward() maps the gradients w.r.t. W_sn (accumulated in the store's gradient bucket by the ordinary backward
    passes) to gradients w.r.t. W, with u and v treated as constants:  dW = (G - <G, W_sn> v u^T) / sigma."""

    def __init__(self, store: ParamStore, variables, seed: int):
        self.store = store
        dev = store.w.device
        gen = torch.Generator().manual_seed(seed + 7919)
        self.entries = []
        for v in variables:
            cols = v.shape[-1]
            rows = v.numel // cols
            u = torch.randn(cols, generator=gen, dtype=torch.float32).to(dev)
            self.entries.append({"var": v, "rows": rows, "cols": cols, "u": u, "u_new": torch.empty_like(u),
                                 "sigma": torch.ones(1, device=dev), "scratch": torch.empty(rows + cols + 4, device=dev),
                                 "dot": torch.empty(1, device=dev)})

    def u_dict(self):
        return {e["var"].name: e["u"].detach().clone() for e in self.entries}

    def load_u(self, us) -> None:
        for e in self.entries:
            if e["var"].name in us:
                e["u"].copy_(torch.as_tensor(us[e["var"].name]).to(device=e["u"].device, dtype=torch.float32).reshape(-1))

    def forward(self, rt, update_u: bool = True) -> None:
        from . import ops
        st = self.store
        st.w_eff.copy_(st.w)                          # biases, BN / CBN parameters, ... are used as they are
        for e in self.entries:
            v = e["var"]
            ops.call.sg_spectral_norm(rt.ctx, ops._p(v.data), e["rows"], e["cols"], ops._p(e["u"]), 1, ops._p(v.eff), ops._p(e["u_new"]),
                                      ops._p(e["sigma"]), ops._p(e["scratch"]))
            if update_u:
                e["u"].copy_(e["u_new"])
        st.version += 1                               # packed filters / the bf16 mirror derive from w_eff

    def backward(self, rt) -> None:
        from . import ops
        for e in self.entries:
            v = e["var"]
            ops.call.sg_spectral_norm_bwd(rt.ctx, ops._p(v.grad), ops._p(v.eff), e["rows"], e["cols"], ops._p(e["u_new"]), ops._p(e["sigma"]),
                                          ops._p(e["scratch"]), ops._p(e["dot"]))
