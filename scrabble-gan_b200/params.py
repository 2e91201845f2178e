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
            ops.call.sg_cast(rt.ctx, ops._p(self.w), ops._p(self.wb), SG_BF16, self.w.numel())
            self.wb_version = self.version
        return self.wb

    @property
    def trainable_variables(self) -> List[Variable]:
        return [v for v in self.vars if v.trainable]

    def zero_grad(self) -> None:
        self.g.zero_()

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
