"""Loss of the reference's CCT ablation (trian_CCT.py:76-84, 213-216): ``softmax_js_loss`` on the device kernel
cmlpl_softmax_js_f32, same signature and preconditions; ``cct_consistency`` is the sum the script forms from it."""
from __future__ import annotations

import torch

from . import ops


class _JSFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, inputs, targets):
        loss, dz = ops.softmax_js(inputs.contiguous(), targets.contiguous(), 1.0, True)
        ctx.save_for_backward(dz)
        return loss

    @staticmethod
    def backward(ctx, g):
        (dz,) = ctx.saved_tensors
        return dz * g, None


def softmax_js_loss(inputs, targets, **_):
    """trian_CCT.py:76-84: 0.5 * (KL(M || softmax(inputs)) + KL(M || targets + 1e-5)) in F.kl_div's 'mean' reduction
    (over every element), M = (softmax(inputs) + targets) / 2; gradient to ``inputs`` only."""
    assert inputs.requires_grad is True and targets.requires_grad is False
    assert inputs.size() == targets.size()
    return _JSFn.apply(inputs, targets.detach())


def cct_consistency(origin_out, aug_out1, aug_out2):
    """trian_CCT.py:211-216: the four JS terms between the clean and the two augmented predictions (targets detached)."""
    ori_t = torch.softmax(origin_out.detach(), dim=1)
    t1 = torch.softmax(aug_out1.detach(), dim=1)
    t2 = torch.softmax(aug_out2.detach(), dim=1)
    return (softmax_js_loss(origin_out, t1) + softmax_js_loss(origin_out, t2) + softmax_js_loss(aug_out1, ori_t) +
            softmax_js_loss(aug_out2, ori_t))
