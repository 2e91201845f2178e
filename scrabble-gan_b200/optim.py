"""Keras-semantics optimizers over flat parameter stores (reference main.py:25-35, data_utils.py:451-468).

    Adam:    lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t);  w -= lr_t * m / (sqrt(v) + 1e-7)
    RMSprop: ms = 0.9 ms + 0.1 g^2;  w -= lr * g / (sqrt(ms) + 1e-7)

apply_gradients(zip(grads, vars)) keeps the reference's call shape.  When the variables are exactly the trainable
variables of one ParamStore (the only way train_step calls it) the update is ONE fused launch over the flat
buffers; otherwise it falls back to one launch per variable."""
from __future__ import annotations

import math
from typing import Dict, Iterable, Tuple

import torch

from . import ops
from .params import ParamStore, Variable


class _FlatState:
    def __init__(self, store: ParamStore, n_slots: int):
        self.slots = [torch.zeros_like(store.w) for _ in range(n_slots)]


class Adam:
    def __init__(self, learning_rate=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-7):
        self.learning_rate, self.beta_1, self.beta_2, self.epsilon = float(learning_rate), float(beta_1), float(beta_2), float(epsilon)
        self.iterations = 0
        self._state: Dict[int, _FlatState] = {}
        self._dev = None            # (step counter int32[1], lr_t float32[1]) on the device: graph-replayable updates
        self._dev_step = None       # host shadow of the device step counter (None = unknown)

    def advance_for_replay(self, rt=None) -> None:
        """Host-side bookkeeping for one CUDA-graph replay of a captured apply_gradients.  The captured sg_adam_prepare
        increments the DEVICE counter; if the host count was changed outside the graph since the last update (restored
        optimizer state, load_state_dict, a roll-back) the device counter is first re-seeded with an explicit launch so
        that the replayed bias-corrected step size uses t = iterations + 1."""
        if self._dev is not None and self._dev_step != self.iterations and rt is not None:
            ops.call.sg_adam_prepare(rt.ctx, ops._p(self._dev[0]), ops._p(self._dev[1]), self.iterations, self.learning_rate,
                                     self.beta_1, self.beta_2)
        self.iterations += 1
        self._dev_step = self.iterations

    def _lr_t(self) -> float:
        t = self.iterations
        return self.learning_rate * math.sqrt(1.0 - self.beta_2 ** t) / (1.0 - self.beta_1 ** t)

    def _slots(self, store: ParamStore) -> _FlatState:
        st = self._state.get(id(store))
        if st is None:
            st = self._state[id(store)] = _FlatState(store, 2)
        return st

    def ensure_state(self, store: ParamStore) -> None:
        """Allocate the (m, v) slots and the device-side step counter for `store` (idempotent)."""
        self._slots(store)
        if self._dev is None:
            rt = store.rt
            self._dev = (torch.zeros(1, device=rt.device, dtype=torch.int32), torch.zeros(1, device=rt.device))

    def apply_gradients(self, grads_and_vars: Iterable[Tuple[torch.Tensor, Variable]]) -> None:
        pairs = list(grads_and_vars)
        if not pairs:
            return
        self.iterations += 1
        lr_t = self._lr_t()
        store = pairs[0][1].store
        rt = store.rt
        st = self._slots(store)
        tv = store.trainable_variables
        fused = len(pairs) == len(tv) and all(v is tv[i] and g.data_ptr() == v.grad.data_ptr() for i, (g, v) in enumerate(pairs))
        mirror_fresh = False
        if fused:
            # keep the bf16 mirror (pack-free tensor-core convs) current in the same pass when it exists and is in sync
            mirror = store.wb if (store.wb is not None and store.wb_version == store.version and store.w_eff is None) else None
            if self._dev is None:
                self._dev = (torch.zeros(1, device=rt.device, dtype=torch.int32), torch.zeros(1, device=rt.device))
            capturing = torch.cuda.is_current_stream_capturing()
            ops.call.sg_adam_prepare(rt.ctx, ops._p(self._dev[0]), ops._p(self._dev[1]), -1 if capturing else self.iterations,
                                     self.learning_rate, self.beta_1, self.beta_2)
            if not capturing:
                self._dev_step = self.iterations
            # one launch per network: the update and the bf16 mirror; with beta_1 == 0 the m slot is not touched.  The gradient
            # is left in place (clear_grad = 0): callers and the parity tests read it after the step
            ops.call.sg_adam_fused(rt.ctx, ops._p(store.w), ops._p(store.g), ops._p(st.slots[0]), ops._p(st.slots[1]), ops._p(mirror),
                                   store.w.numel(), ops._p(self._dev[1]), self.beta_1, self.beta_2, self.epsilon, 0)
            mirror_fresh = mirror is not None
        else:
            for g, v in pairs:
                assert v.store is store, "one optimizer serves one network"
                sl = slice(v.offset, v.offset + v.numel)
                gg = g if isinstance(g, torch.Tensor) else torch.as_tensor(g)
                gg = gg.to(device=rt.device, dtype=torch.float32).reshape(-1).contiguous()
                ops.adam_(rt, store.w[sl], gg, st.slots[0][sl], st.slots[1][sl], lr_t, self.beta_1, self.beta_2, self.epsilon)
        store.version += 1
        if mirror_fresh:
            store.wb_version = store.version

    def state_dict(self):
        return {"iterations": self.iterations, "slots": {k: [s.clone() for s in v.slots] for k, v in self._state.items()}}

    def load_state_dict(self, sd) -> None:
        """Restore `iterations` and the (m, v) slots saved by state_dict(); the device step counter is re-seeded on the next
        update (eager: explicit t; replayed: advance_for_replay notices the mismatch)."""
        self.iterations = int(sd["iterations"])
        for k, slots in sd.get("slots", {}).items():
            st = self._state.get(k)
            if st is None:
                continue
            for t, src in zip(st.slots, slots):
                t.copy_(src)
        self._dev_step = None


class RMSprop:
    def __init__(self, learning_rate=0.001, rho=0.9, epsilon=1e-7):
        self.learning_rate, self.rho, self.epsilon = float(learning_rate), float(rho), float(epsilon)
        self.iterations = 0
        self._state: Dict[int, _FlatState] = {}

    def apply_gradients(self, grads_and_vars) -> None:
        pairs = list(grads_and_vars)
        if not pairs:
            return
        self.iterations += 1
        store = pairs[0][1].store
        rt = store.rt
        st = self._state.get(id(store))
        if st is None:
            st = self._state[id(store)] = _FlatState(store, 1)
        tv = store.trainable_variables
        fused = len(pairs) == len(tv) and all(v is tv[i] and g.data_ptr() == v.grad.data_ptr() for i, (g, v) in enumerate(pairs))
        if fused:
            ops.rmsprop_(rt, store.w, store.g, st.slots[0], self.learning_rate, self.rho, self.epsilon)
        else:
            for g, v in pairs:
                sl = slice(v.offset, v.offset + v.numel)
                gg = g.to(device=rt.device, dtype=torch.float32).reshape(-1).contiguous()
                ops.rmsprop_(rt, store.w[sl], gg, st.slots[0][sl], self.learning_rate, self.rho, self.epsilon)
        store.version += 1


def setup_optimizer(g_lr, d_lr, r_lr, w_lr, beta_1, beta_2, loss_fn, disc_iters, apply_gradient_balance, rmsprop):
    """Same signature and return tuple as the reference's gin-configurable setup_optimizer (main.py:25-35)."""
    generator_optimizer = Adam(learning_rate=g_lr, beta_1=beta_1, beta_2=beta_2)
    discriminator_optimizer = Adam(learning_rate=d_lr, beta_1=beta_1, beta_2=beta_2)
    if rmsprop:
        recognizer_optimizer = RMSprop(learning_rate=r_lr)
    else:
        recognizer_optimizer = Adam(learning_rate=r_lr, beta_1=beta_1, beta_2=beta_2)
    stylepromoter_optimizer = Adam(learning_rate=w_lr, beta_1=beta_1, beta_2=beta_2)
    return (generator_optimizer, discriminator_optimizer, recognizer_optimizer, stylepromoter_optimizer, loss_fn, disc_iters,
            apply_gradient_balance)
