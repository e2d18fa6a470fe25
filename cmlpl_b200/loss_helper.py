"""The callables of the reference's loss_helper.py (U2PL-derived; imported by trian_CPS.py /
trian_CCT.py, never called by any script of the reference) with the same signatures, return
conventions and side effects (in-place mutation of ``target`` and of the memory bank lists).

On the per-pixel ``[N, C]`` shapes of this path ``compute_unsupervised_loss`` runs on the
libcmlpl_sm100.so kernels (row entropy, masked cross entropy forward+backward); the percentile is
taken from a device sort and interpolated like ``np.percentile`` entirely on the device -- no host
synchronisation (the reference does a full D2H + numpy sort there).  The segmentation-style 4-D criteria (Criterion*, Ohem*) and the ReCo memory-bank
loss are bookkeeping-heavy and shape-incompatible with the hot path; they are kept as thin device
tensor code that routes every cross entropy through the same kernel.
"""
from __future__ import annotations

import numpy as np
import torch
from torch import nn

from . import losses, ops

IGNORE = 255


# ------------------------------------------------------------------ helpers
def _masked_ce(logits2d, target1d, ignore_index=IGNORE, class_weight=None):
    """CrossEntropyLoss(ignore_index, weight) with mean reduction on [n, C] logits, through
    cmlpl_ce_fwd_bwd_f32: mean over the kept rows (weighted by class_weight[target] if given)."""
    keep = target1d != ignore_index
    w = keep.float() if class_weight is None else torch.where(keep, class_weight[target1d.clamp(max=len(class_weight) - 1)], torch.zeros((), device=logits2d.device))
    denom = w.sum()
    n = logits2d.size(0)
    # the kernel returns mean_i(w_i * ce_i) over all n rows; rescale to sum/denom
    loss = losses._CEFn.apply(logits2d.contiguous(), target1d.contiguous(), None, w.contiguous())
    return loss * (n / denom)


def _flatten_nchw(pred):
    b, c, h, w = pred.shape
    return pred.permute(0, 2, 3, 1).reshape(b * h * w, c)


def _percentile_linear(values: torch.Tensor, valid: torch.Tensor, q: float) -> torch.Tensor:
    """np.percentile(values[valid], q) (linear interpolation) as a 0-dim DEVICE tensor, with no host synchronisation:
    invalid entries sort to the end as +inf, the interpolation position comes from the device-side count.  The lerp
    is numpy's (``a + (b-a)*t``, switched to ``b - (b-a)*(1-t)`` for t >= 0.5) in float64, rounded to the input dtype."""
    srt = torch.sort(torch.where(valid, values, torch.full_like(values, float("inf"))))[0]
    n = valid.sum()
    pos = (n - 1).double() * (q / 100.0)
    lo = torch.floor(pos).long().clamp(min=0)
    hi = torch.minimum(lo + 1, (n - 1).clamp(min=0))
    a, b = srt[lo].double(), srt[hi].double()
    t = pos - lo.double()
    out = torch.where(t >= 0.5, b - (b - a) * (1 - t), a + (b - a) * t)
    return out.to(values.dtype)


# ------------------------------------------------------------------ loss_helper.py:19-36
def dequeue_and_enqueue(keys, queue, queue_ptr, queue_size):
    """Append ``keys`` to the single-tensor list ``queue`` keeping the newest ``queue_size`` rows;
    ``queue_ptr[0]`` follows the reference's pointer rule.  Returns the number of keys."""
    keys = keys.detach().clone().cpu()
    count = keys.shape[0]
    queue[0] = torch.cat((queue[0], keys), dim=0)
    if queue[0].shape[0] >= queue_size:
        queue[0] = queue[0][-queue_size:, :]
        queue_ptr[0] = queue_size
    else:
        queue_ptr[0] = (int(queue_ptr) + count) % queue_size
    return count


# ------------------------------------------------------------------ loss_helper.py:242-261
def compute_unsupervised_loss(predict, target, percent, pred_teacher):
    """Drop the ``percent``-th-percentile-and-above entropy pixels of the teacher (``target`` <- 255 in
    place), then weight * CE(predict, target, ignore=255) with weight = N / #kept."""
    batch_size, _ = predict.shape
    with torch.no_grad():
        entropy = ops.softmax_entropy(pred_teacher.detach().contiguous().float(), 1e-10)
        valid = target != IGNORE
        thresh = _percentile_linear(entropy, valid, percent)            # device scalar: the reference syncs here
        target.masked_fill_(entropy.ge(thresh) & valid, IGNORE)
        weight = batch_size / torch.sum(target != IGNORE)
    return weight * _masked_ce(predict, target)


# ------------------------------------------------------------------ loss_helper.py:222-239
def compute_rce_loss(predict, target):
    """Reverse CE on [b, C, h, w] predictions: -sum_c softmax(p)_c log(clamp(onehot_c, 1e-4, 1)),
    averaged over target != 255."""
    prob = torch.softmax(predict, dim=1)
    keep = target != IGNORE
    with torch.no_grad():
        hot = torch.zeros_like(prob).scatter_(1, torch.where(keep, target, torch.zeros_like(target)).unsqueeze(1), 1.0)
        log_label = torch.log(hot.clamp(min=1e-4, max=1.0))
    rce = -(prob * log_label).sum(1) * keep
    return rce.sum() / keep.sum()


# ------------------------------------------------------------------ loss_helper.py:39-219
def compute_contra_memobank_loss(rep, label_l, label_u, prob_l, prob_u, low_mask, high_mask, memobank, queue_prtlis,
                                 queue_size, rep_teacher, momentum_prototype=None, i_iter=0):
    """ReCo-style contrastive loss against per-class negative banks (constants loss_helper.py:56-61).
    Per class: anchors = low-entropy representations with prob > 0.3, positive = mean teacher
    representation of the class, 50 negatives per anchor sampled from the class bank, cosine logits / 0.5,
    CE towards index 0.  Returns (new_keys, loss) or (prototype, new_keys, loss) like the reference,
    including its indexing of the per-class lists by *valid-class order* (loss_helper.py:150-171)."""
    pos_thresh, neg_thresh, low_rank, high_rank, temp, num_queries, num_negatives = 0.3, 1, 3, 9, 0.5, 256, 50
    num_feat, num_labeled, num_segments = rep.shape[1], label_l.shape[0], label_l.shape[1]
    labels_all = torch.cat((label_l, label_u), dim=0)
    low_valid, high_valid = labels_all * low_mask, labels_all * high_mask
    rank_l = torch.sort(prob_l, 1, True)[1]
    rank_u = torch.sort(prob_u, 1, True)[1]
    prob = torch.cat((prob_l, prob_u), dim=0)
    anchors_by_class, protos, counts, valid_classes, new_keys = [], [], [], [], []
    for cls in range(num_segments):
        low_c, high_c = low_valid[:, cls].bool(), high_valid[:, cls].bool()
        anchors_by_class.append(rep[(prob[:, cls] > pos_thresh) & low_c])
        protos.append(torch.mean(rep_teacher[low_c].detach(), dim=0, keepdim=True))
        in_rank_u = rank_u[:, low_rank:high_rank].eq(cls).sum(1).bool()
        in_rank_l = rank_l[:, :low_rank].eq(cls).sum(1).bool() & (label_l[:, cls] == 0)
        negatives = ((prob[:, cls] < neg_thresh) & high_c) & torch.cat((in_rank_l, in_rank_u), dim=0)
        new_keys.append(dequeue_and_enqueue(rep_teacher[negatives].detach(), memobank[cls], queue_prtlis[cls], queue_size[cls]))
        if low_c.sum() > 0:
            counts.append(int(low_c.sum().item()))
            valid_classes.append(cls)
    zero = torch.tensor(0.0, device=rep.device) * rep.sum()
    if len(counts) <= 1:
        return (new_keys, zero) if momentum_prototype is None else (momentum_prototype, new_keys, zero)
    proto_all = torch.cat(protos)
    prototype = torch.zeros((rank_l.shape[-1], num_queries, 1, num_feat), device=rep.device)
    reco = torch.zeros((), device=rep.device)
    for j, cls in enumerate(valid_classes):
        bank = memobank[cls][0]
        if len(anchors_by_class[j]) == 0 or bank.shape[0] == 0:     # [j], not [cls]: reference indexing
            reco = reco + 0 * rep.sum()
            continue
        pick = torch.randint(len(anchors_by_class[j]), size=(num_queries,))
        anchor = anchors_by_class[j][pick.to(rep.device)]
        with torch.no_grad():
            neg = bank.to(rep.device)
            neg = neg[torch.randint(len(neg), size=(num_queries * num_negatives,)).to(rep.device)]
            neg = neg.reshape(num_queries, num_negatives, num_feat)
            pos = proto_all[j].view(1, 1, -1).repeat(num_queries, 1, 1)
            if momentum_prototype is not None:
                if not (momentum_prototype == 0).all():
                    decay = min(1 - 1 / i_iter, 0.999)
                    pos = (1 - decay) * pos + decay * momentum_prototype[cls]
                prototype[cls] = pos.clone()
            cand = torch.cat((pos, neg), dim=1)
        logits = torch.cosine_similarity(anchor.unsqueeze(1), cand, dim=2) / temp
        reco = reco + losses.cross_entropy(logits, torch.zeros(num_queries, dtype=torch.long, device=rep.device))
    out = reco / len(counts)
    return (new_keys, out) if momentum_prototype is None else (prototype, new_keys, out)


# ------------------------------------------------------------------ loss_helper.py:264-557 (segmentation-style criteria)
_CITYSCAPES_BINARY = [0, 0, 0, 1, 1, 1, 1, 0, 0, 1, 0, 0, 1, 0, 1, 0, 1, 1, 1]
_CITYSCAPES_OHEM = [0.8373, 0.918, 0.866, 1.0345, 1.0166, 0.9969, 0.9754, 1.0489, 0.8786, 1.0023, 0.9539, 0.9843,
                    1.1116, 0.9037, 1.0865, 1.0955, 1.0865, 1.1529, 1.0507]


def _seg_ce(pred, target, ignore_index, class_weight=None):
    cw = None if class_weight is None else torch.tensor(class_weight, dtype=torch.float32, device=pred.device)
    return _masked_ce(_flatten_nchw(pred), target.reshape(-1), ignore_index, cw)


def _check_same_hw(preds, target, aux):
    h, w = target.size(1), target.size(2)
    for p in (preds if aux else (preds,)):
        assert p.size(2) == h and p.size(3) == w
    if aux:
        assert len(preds) == 2


class Criterion(nn.Module):
    """CE (+ aux-head CE * aux_weight; + a 19-class binary-weighted CE when use_weight)."""

    def __init__(self, aux_weight, ignore_index=IGNORE, use_weight=False):
        super().__init__()
        self._aux_weight, self._ignore_index, self.use_weight = aux_weight, ignore_index, use_weight

    def _main(self, pred, target):
        loss = _seg_ce(pred, target, self._ignore_index)
        if self.use_weight:
            loss = loss + _seg_ce(pred, target, self._ignore_index, _CITYSCAPES_BINARY)
        return loss

    def forward(self, preds, target):
        aux = self._aux_weight > 0
        _check_same_hw(preds, target, aux)
        if not aux:
            return _seg_ce(preds, target, self._ignore_index)
        return self._main(preds[0], target) + self._aux_weight * _seg_ce(preds[1], target, self._ignore_index)


class OhemCrossEntropy2dTensor(nn.Module):
    """Online hard example mining on device: keep pixels whose true-class probability is <= max(thresh,
    the min_kept-th smallest), ignore the rest, then CE."""

    def __init__(self, ignore_index=IGNORE, thresh=0.7, min_kept=256, use_weight=False, reduce=False):
        super().__init__()
        self.ignore_index, self.thresh, self.min_kept = ignore_index, float(thresh), int(min_kept)
        self.class_weight = _CITYSCAPES_OHEM if use_weight else None
        self.reduce = reduce and not use_weight

    def forward(self, pred, target):
        b, c, h, w = pred.size()
        flat = target.reshape(-1).clone()
        valid = flat.ne(self.ignore_index)
        num_valid = int(valid.sum())
        if self.min_kept <= num_valid and num_valid > 0:
            prob = torch.softmax(pred, dim=1).transpose(0, 1).reshape(c, -1)
            true_p = prob.gather(0, (flat * valid.long()).unsqueeze(0)).squeeze(0).masked_fill(~valid, 1.0)
            threshold = self.thresh
            if self.min_kept > 0:
                kth = torch.sort(true_p)[0][min(true_p.numel(), self.min_kept) - 1]
                threshold = max(threshold, float(kth))
                valid = valid & true_p.le(threshold)
        flat = flat.masked_fill(~valid, self.ignore_index)
        if self.reduce:
            return torch.nn.functional.cross_entropy(pred, flat.view(b, h, w), ignore_index=self.ignore_index, reduction="none")
        return _seg_ce(pred, flat.view(b, h, w), self.ignore_index, self.class_weight)


class CriterionOhem(nn.Module):
    def __init__(self, aux_weight, thresh=0.7, min_kept=100000, ignore_index=IGNORE, use_weight=False):
        super().__init__()
        self._aux_weight = aux_weight
        self._criterion1 = OhemCrossEntropy2dTensor(ignore_index, thresh, min_kept, use_weight)
        self._criterion2 = OhemCrossEntropy2dTensor(ignore_index, thresh, min_kept)

    def forward(self, preds, target):
        aux = self._aux_weight > 0
        _check_same_hw(preds, target, aux)
        if not aux:
            return self._criterion1(preds, target)
        return self._criterion1(preds[0], target) + self._aux_weight * self._criterion2(preds[1], target)


class OhemCrossEntropy2d(nn.Module):
    """Host-thresholded OHEM (the reference zooms predictions down by ``factor`` with scipy before
    picking the threshold, loss_helper.py:400-430)."""

    def __init__(self, ignore_label=IGNORE, thresh=0.7, min_kept=100000, factor=8):
        super().__init__()
        self.ignore_label, self.thresh, self.min_kept, self.factor = ignore_label, float(thresh), int(min_kept), factor

    def find_threshold(self, np_predict, np_target):
        import scipy.ndimage as nd
        f = self.factor
        predict = nd.zoom(np_predict, (1.0, 1.0, 1.0 / f, 1.0 / f), order=1)
        target = nd.zoom(np_target, (1.0, 1.0 / f, 1.0 / f), order=0)
        c = predict.shape[1]
        min_kept = self.min_kept // (f * f)
        labels = target.ravel().astype(np.int32)
        probs = np.rollaxis(predict, 1).reshape((c, -1))
        valid = labels != self.ignore_label
        if min_kept >= valid.sum():
            return 1.0
        threshold = self.thresh
        if valid.sum() > 0 and min_kept > 0:
            true_p = probs[:, valid][labels[valid], np.arange(valid.sum(), dtype=np.int32)]
            kth = np.partition(true_p, min(len(true_p), min_kept) - 1)[min(len(true_p), min_kept) - 1]
            threshold = max(threshold, kth)
        return threshold

    def generate_new_target(self, predict, target):
        np_predict, np_target = predict.data.cpu().numpy(), target.data.cpu().numpy()
        c = np_predict.shape[1]
        threshold = self.find_threshold(np_predict, np_target)
        labels = np_target.ravel().astype(np.int32)
        probs = np.rollaxis(np_predict, 1).reshape((c, -1))
        valid_inds = np.where(labels != self.ignore_label)[0]
        if len(valid_inds) > 0:
            true_p = probs[:, valid_inds][labels[valid_inds], np.arange(len(valid_inds), dtype=np.int32)]
            valid_inds = valid_inds[true_p <= threshold]
        kept = labels[valid_inds].copy()
        labels.fill(self.ignore_label)
        labels[valid_inds] = kept
        return torch.from_numpy(labels.reshape(target.size())).long().to(target.device)

    def forward(self, predict, target, weight=None):
        assert not target.requires_grad
        new_target = self.generate_new_target(torch.softmax(predict, 1), target)
        return _seg_ce(predict, new_target, self.ignore_label)


def get_criterion(cfg):
    """loss_helper.py:264-281."""
    crit = cfg["criterion"]
    aux_weight = cfg["net"]["aux_loss"]["loss_weight"] if cfg["net"].get("aux_loss", False) else 0
    ignore_index = cfg["dataset"]["ignore_label"]
    cls = CriterionOhem if crit["type"] == "ohem" else Criterion
    return cls(aux_weight, ignore_index=ignore_index, **crit["kwargs"])
