"""Stub of the module the reference's ablation scripts import but do not ship (``from regularizer import
Distribution_Loss``, trian_CPS.py:11; constructed at :163 and never called).  Present so that the ablation entry points
import; calling it raises, because the reference defines no arithmetic for it."""
from torch import nn


class Distribution_Loss(nn.Module):
    def __init__(self, loss='mmd', *args, **kwargs):
        super().__init__()
        self.loss = loss

    def forward(self, *args, **kwargs):
        raise NotImplementedError("regularizer.Distribution_Loss is not part of liuli33/CMLPL (trian_CPS.py:11 imports a "
                                  "module the repository does not contain; the script never calls it)")
