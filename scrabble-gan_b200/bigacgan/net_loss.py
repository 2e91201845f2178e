"""hinge / not_saturating with the reference's signatures (src/bigacgan/net_loss.py:38-54 and :4-35).

Both take (B,1) logit tensors living on the GPU (torch tensors or any DLPack producer) and return the reference's
7-tuple (d_loss, d_loss_real, d_loss_fake, g_loss, s_loss, s_loss_1, s_loss_2) of (B,1) device tensors, computed by
one libsgan launch.  train_step recognises these two functions and uses the fused loss + gradient-balance kernels
(which also emit the per-sample upstream gradients); any other callable is rejected there."""
from __future__ import annotations

import torch

from .._abi import SG_LOSS_HINGE, SG_LOSS_NOT_SATURATING, call, from_dlpack
from ..ops import _p
from ..runtime import get_runtime


def _terms(kind, d_real, d_fake, s_a, s_b, s_c=None):
    rt = get_runtime()
    ts = [None if t is None else from_dlpack(t).to(device=rt.device, dtype=torch.float32).reshape(-1).contiguous()
          for t in (d_real, d_fake, s_a, s_b, s_c)]
    b = ts[0].numel()
    out = rt.empty((7, b))
    call.sg_loss_terms(rt.ctx, kind, _p(ts[0]), _p(ts[1]), _p(ts[2]), _p(ts[3]), _p(ts[4]), b, _p(out))
    return tuple(out[i].view(b, 1) for i in range(7))


def hinge(d_real_logits, d_fake_logits, s_real_logits, s_fake_logits, *ignored):
    """relu(1 - d_real) + relu(1 + d_fake), same for the style promoter, g = -(d_fake + s_fake).
    The reference's train_step passes a 5th positional argument that hinge does not accept; extra
    positional arguments are therefore accepted and ignored."""
    return _terms(SG_LOSS_HINGE, d_real_logits, d_fake_logits, s_real_logits, s_fake_logits, None)


def not_saturating(d_real_logits, d_fake_logits, s_styleimgs_logits, s_trainingimgs_logits, s_fake_logits):
    """sigmoid cross-entropy losses, positionally bug-compatible with the reference (its train_step passes
    (d_real, d_fake, s_real, s_fake, s_real_real_imgs), so the generator term uses the 5th slot)."""
    return _terms(SG_LOSS_NOT_SATURATING, d_real_logits, d_fake_logits, s_styleimgs_logits, s_trainingimgs_logits, s_fake_logits)


hinge.sg_kind = SG_LOSS_HINGE
not_saturating.sg_kind = SG_LOSS_NOT_SATURATING
