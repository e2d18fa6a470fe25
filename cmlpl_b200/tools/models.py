"""Network constructors of the reference's tools/models.py, backed by libcmlpl_sm100.so.

``BaseNet2`` keeps the reference's constructor, parameter names/shapes (so state dicts
interchange, models.py:102-127), default initialisation order (same seed -> same weights)
and ``forward(x, y) -> (logits, l2norm(relu(feat_spe(y))))`` (models.py:130-152); the
arithmetic runs in the fp32 CUDA kernels of cmlpl_b200/csrc/fp32_ops.cu through a custom
autograd function (forward and backward are both ours; torch is used for memory, the
channel concat and the dropout mask only).
"""
from __future__ import annotations

import torch
from torch import nn

from .. import ops

CONV_IN = 60      # conv0 is hard-wired to 60 PCA channels (models.py:102)
CONV_CH = 64
N_FC1 = 1024


class _BaseNet2Fn(torch.autograd.Function):
    """conv0 -> [conv3x3 + res + relu -> avgpool2] x2 -> flatten || relu(linear) -> classifier."""

    @staticmethod
    def forward(ctx, x, y, w0, b0, w1, b1, w2, b2, ws, bs, wc, bc, mask):
        x = x.contiguous()
        y = y.contiguous()
        a0 = ops.conv2d(x, w0, b0)                               # models.py:132
        a1 = ops.conv2d(a0, w1, b1, res=a0, relu=True)           # :133-135
        p1 = ops.avgpool2(a1)                                    # :136
        a2 = ops.conv2d(p1, w2, b2, res=p1, relu=True)           # :137-139
        p2 = ops.avgpool2(a2)                                    # :140
        h = ops.sgemm(y, ws, transB=True, bias=bs, act=1)        # :142-143
        cat = torch.cat([p2.reshape(p2.size(0), -1), h], 1)      # :141,144
        feat, norm = ops.l2norm(h)                               # :145-146
        cat_d = cat * mask if mask is not None else cat          # :147-148 (mask = bernoulli/(1-p))
        logits = ops.sgemm(cat_d, wc, transB=True, bias=bc)      # :150
        ctx.save_for_backward(x, y, w1, w2, ws, wc, a0, a1, p1, a2, h, cat_d, feat, norm,
                              mask if mask is not None else torch.empty(0, device=x.device))
        ctx.has_mask = mask is not None
        ctx.p2_shape = p2.shape
        return logits, feat

    @staticmethod
    def backward(ctx, dlogits, dfeat):
        x, y, w1, w2, ws, wc, a0, a1, p1, a2, h, cat_d, feat, norm, mask = ctx.saved_tensors
        dlogits = dlogits.contiguous()
        b = dlogits.size(0)
        # classifier
        dwc = ops.sgemm(dlogits, cat_d, transA=True)             # [C, 2624]
        dbc = ops.colsum(dlogits)
        dcat = ops.sgemm(dlogits, wc)                            # [b, 2624]
        if ctx.has_mask:
            dcat = dcat * mask
        nconv = dcat.size(1) - N_FC1
        dh = dcat[:, nconv:].contiguous()
        if dfeat is not None:
            dh = dh + ops.l2norm_bwd(feat, norm, dfeat.contiguous())
        dhp = ops.relu_bwd(h, dh)
        dws = ops.sgemm(dhp, y, transA=True)                     # [1024, B]
        dbs = ops.colsum(dhp)
        # conv trunk
        dp2 = dcat[:, :nconv].contiguous().view(ctx.p2_shape)
        da2 = ops.avgpool2_bwd(dp2, a2.size(2), a2.size(3))
        dz2 = ops.relu_bwd(a2, da2)
        dw2, db2 = ops.conv2d_wgrad(p1, dz2, 3)
        dp1 = ops.conv2d_dgrad(dz2, w2, res=dz2)                 # conv path + residual path
        da1 = ops.avgpool2_bwd(dp1, a1.size(2), a1.size(3))
        dz1 = ops.relu_bwd(a1, da1)
        dw1, db1 = ops.conv2d_wgrad(a0, dz1, 3)
        da0 = ops.conv2d_dgrad(dz1, w1, res=dz1)
        dw0, db0 = ops.conv2d_wgrad(x, da0, 1)
        return None, None, dw0, db0, dw1, db1, dw2, db2, dws, dbs, dwc, dbc, None


class ContrastiveLoss(nn.Module):
    """models.py:14-39 (SimCLR NT-Xent over the [2bs, 2bs] cosine matrix) through cmlpl_ntxent_f32:
    normalise -> similarity GEMM -> masked log-sum-exp, gradient in the same pass (never the
    O(n^2 d) broadcast the reference allocates)."""

    def __init__(self, batch_size, device='cuda', temperature=0.5):
        super().__init__()
        self.batch_size = batch_size
        self.register_buffer("temperature", torch.tensor(temperature).to(device))
        self.register_buffer("negatives_mask",
                             (~torch.eye(batch_size * 2, batch_size * 2, dtype=bool).to(device)).float())
        self._t = float(temperature)

    def forward(self, emb_i, emb_j):
        from ..losses import nt_xent
        if emb_i.size(0) != self.batch_size:
            raise ValueError("ContrastiveLoss was built for batch_size=%d" % self.batch_size)
        return nt_xent(emb_i, emb_j, self._t)


class Normalize(nn.Module):
    """models.py:81-90: x / ||x||_p along dim 1, no epsilon (power 2 only on device)."""

    def __init__(self, power=2):
        super().__init__()
        self.power = power

    def forward(self, x):
        if self.power != 2:
            raise NotImplementedError("only the L2 form is used on the hot path (models.py:128)")
        return _L2NormFn.apply(x)


class _L2NormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        y, norm = ops.l2norm(x.contiguous())
        ctx.save_for_backward(y, norm)
        return y

    @staticmethod
    def backward(ctx, dy):
        y, norm = ctx.saved_tensors
        return ops.l2norm_bwd(y, norm, dy.contiguous())


class BaseNet2(nn.Module):
    def __init__(self, num_features=103, dropout=0, num_classes=0, w=20):
        """Reference signature (models.py:98) plus ``w``: the reference hard-wires the classifier to 2624 inputs, i.e.
        w = 20 (models.py:127).  ``w=11`` builds the odd-window variant BASELINE configs[4] names -- windows as in
        ExtractPatches_for_base (hyper_tools.py:300-317), pooled 11 -> 5 -> 2, classifier over 64*2*2 + 1024 = 1280
        inputs -- a documented extension; every other layer is unchanged."""
        super().__init__()
        self.w = w
        # creation order == reference (models.py:102-127) so a given torch seed yields the same weights
        self.conv0 = nn.Conv2d(CONV_IN, CONV_CH, kernel_size=1, stride=1, bias=True)
        self.conv1 = nn.Conv2d(CONV_CH, CONV_CH, kernel_size=3, stride=1, padding=1, bias=True)
        self.conv2 = nn.Conv2d(CONV_CH, CONV_CH, kernel_size=3, stride=1, padding=1, bias=True)
        self.num_features = num_features
        self.dropout = dropout
        self.drop = nn.Dropout(self.dropout)
        self.num_classes = num_classes
        self.feat_spe = nn.Linear(num_features, N_FC1)
        # present in the reference's state dict but never used by forward (models.py:122-126)
        self.feat_ss = nn.Linear(N_FC1, 256)
        self.feat_ss2 = nn.Linear(N_FC1, 64)
        self.feat_ss3 = nn.Linear(256, 64)
        self.classifier = nn.Linear(CONV_CH * ((w // 2) // 2) ** 2 + N_FC1, num_classes)
        self.l2norm = Normalize(2)
        self._packed = None   # (version key, packed weights) cache for scene inference

    def forward(self, x, y, dropout_mask=None):
        """x f32 [b,60,w,w] (contiguous NCHW), y f32 [b,num_features] -> (logits, feat).

        ``dropout_mask`` (already scaled by 1/(1-p)) lets a test harness inject the mask;
        otherwise inverted dropout with p=self.dropout is drawn when training (models.py:147-148)."""
        mask = dropout_mask
        if mask is None and self.dropout > 0 and self.training:
            mask = self.drop(torch.ones((x.size(0), self.classifier.in_features), dtype=x.dtype, device=x.device))
        return _BaseNet2Fn.apply(x, y, self.conv0.weight, self.conv0.bias, self.conv1.weight, self.conv1.bias,
                                 self.conv2.weight, self.conv2.bias, self.feat_spe.weight, self.feat_spe.bias,
                                 self.classifier.weight, self.classifier.bias, mask)

    # ---- scene inference support (used by tools.hyper_tools.test_whole)
    def packed_weights(self, w: int = 20):
        key = tuple((p.data_ptr(), p._version) for p in self.parameters()) + (w,)
        if self._packed is None or self._packed[0] != key:
            sd = {k: v for k, v in self.state_dict().items()}
            self._packed = (key, ops.pack_basenet2(sd, self.num_features, self.num_classes, w))
        return self._packed[1]
