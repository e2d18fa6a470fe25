"""Losses of the reference's CPS (cross pseudo supervision) ablation, trian_CPS.py:232-250, on the device CE kernel:
each net is supervised by the other's hard pseudo-labels on the unlabelled rows (weight 0.1)."""
from __future__ import annotations

import torch

from . import losses, ops
from .regularizer import Distribution_Loss  # noqa: F401  (trian_CPS.py:11)


def cps_losses(out_b, out_e, Y_train):
    """out_b / out_e: logits of Base / Base1 on [labelled ; unlabelled] rows (trian_CPS.py:211-228), Y_train i64 [bs].
    Returns (total_loss, total_loss1, parts) with parts = (cls, cls1, con, con1) exactly as trian_CPS.py:232-247."""
    bs = Y_train.size(0)
    labeled_output, un_b_output = out_b[:bs], out_b[bs:]
    labeled_output1, un_e_output = out_e[:bs], out_e[bs:]
    cls = losses.cross_entropy(labeled_output, Y_train)                                   # :232
    cls1 = losses.cross_entropy(labeled_output1, Y_train)                                 # :233
    prd1 = ops.argmax_u8(un_b_output.detach().contiguous()).to(torch.int64)               # :236
    prd2 = ops.argmax_u8(un_e_output.detach().contiguous()).to(torch.int64)               # :237
    con = losses.cross_entropy(un_b_output, prd2)                                         # :239,241
    con1 = losses.cross_entropy(un_e_output, prd1)                                        # :240,242
    return cls + 0.1 * con, cls1 + 0.1 * con1, (cls, cls1, con, con1)                     # :243,246
