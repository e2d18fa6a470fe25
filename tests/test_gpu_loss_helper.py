"""loss_helper callables against outputs of the reference's own loss_helper.py
(tests/golden/loss_helper.npz, metrics_losses.npz; oracle/make_golden.py).  -m gpu."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _z(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def test_compute_unsupervised_loss_matches_reference(dev, golden_dir):
    from cmlpl_b200 import loss_helper as LH
    z = _z(golden_dir, "metrics_losses.npz")
    predict = torch.from_numpy(z["us_predict"]).to(dev).requires_grad_(True)
    target = torch.from_numpy(z["us_target"]).to(dev)
    loss = LH.compute_unsupervised_loss(predict, target, 80, torch.from_numpy(z["us_teacher"]).to(dev))
    assert abs(float(loss) - float(z["us_loss"])) < 1e-5
    assert np.array_equal(target.cpu().numpy(), z["us_target_after"])        # in-place mutation like the reference
    loss.backward()
    # gradient against torch's own CE on the mutated target
    p = torch.from_numpy(z["us_predict"]).requires_grad_(True)
    t = torch.from_numpy(z["us_target_after"])
    w = 200 / (t != 255).sum()
    (w * torch.nn.functional.cross_entropy(p, t, ignore_index=255)).backward()
    assert np.abs(predict.grad.cpu().numpy() - p.grad.numpy()).max() <= 1e-5 * np.abs(p.grad.numpy()).max()


def test_segmentation_criteria_match_reference(dev, golden_dir):
    from cmlpl_b200 import loss_helper as LH
    z = _z(golden_dir, "loss_helper.npz")
    pred = torch.from_numpy(z["seg_pred"]).to(dev); aux = torch.from_numpy(z["seg_aux"]).to(dev)
    tgt = torch.from_numpy(z["seg_tgt"]).to(dev)
    close = lambda a, k: abs(float(a) - float(z[k])) < 2e-5 * max(1.0, abs(float(z[k])))
    assert close(LH.Criterion(0)(pred, tgt), "crit_plain")
    assert close(LH.Criterion(0.4)((pred, aux), tgt), "crit_aux")
    assert close(LH.Criterion(0.4, use_weight=True)((pred, aux), tgt), "crit_weight")
    assert close(LH.OhemCrossEntropy2dTensor(255, 0.7, 40)(pred, tgt.clone()), "ohem_tensor")
    assert close(LH.OhemCrossEntropy2dTensor(255, 0.05, 30, use_weight=True)(pred, tgt.clone()), "ohem_tensor_w")
    assert close(LH.CriterionOhem(0.4, thresh=0.7, min_kept=50)((pred, aux), tgt.clone()), "crit_ohem")
    cfg = {"criterion": {"type": "ohem", "kwargs": {"thresh": 0.7, "min_kept": 50}},
           "net": {"aux_loss": {"loss_weight": 0.4}}, "dataset": {"ignore_label": 255}}
    assert close(LH.get_criterion(cfg)((pred, aux), tgt.clone()), "get_criterion")
    assert close(LH.OhemCrossEntropy2d(255, 0.7, 160, factor=2)(pred, tgt.clone()), "ohem_host")
    assert close(LH.compute_rce_loss(pred, tgt.clone()), "rce")


def test_dequeue_and_enqueue_matches_reference(golden_dir):
    from cmlpl_b200 import loss_helper as LH
    z = _z(golden_dir, "loss_helper.npz")
    g = torch.Generator().manual_seed(21)
    # replay the generator up to the keys (same draws as make_golden.gold_loss_helper)
    torch.randn(2, 19, 12, 10, generator=g); torch.randn(2, 19, 12, 10, generator=g); torch.randint(0, 19, (2, 12, 10), generator=g)
    q, ptr, ptrs = [torch.zeros(0, 8)], torch.zeros(1, dtype=torch.long), []
    for i, n in enumerate((5, 9, 4)):
        LH.dequeue_and_enqueue(torch.randn(n, 8, generator=g), q, ptr, 12)
        ptrs.append(int(ptr[0]))
        if i == 1:
            assert np.array_equal(q[0].numpy(), z["dq_mid"])
    assert np.array_equal(q[0].numpy(), z["dq_final"]) and ptrs == list(z["dq_ptrs"])


def test_contra_memobank_loss_matches_reference(dev, golden_dir):
    from cmlpl_b200 import loss_helper as LH
    z = _z(golden_dir, "loss_helper.npz")
    C, D = 6, 32
    t = lambda k: torch.from_numpy(z[k]).to(dev)

    def banks():
        gg = torch.Generator().manual_seed(5)
        return [[torch.randn(7, D, generator=gg)] for _ in range(C)], [torch.zeros(1, dtype=torch.long) for _ in range(C)]

    mb, ptrs = banks()
    torch.manual_seed(77)
    rep = t("mb_rep").requires_grad_(True)
    keys, loss = LH.compute_contra_memobank_loss(rep, t("mb_label_l"), t("mb_label_u"), t("mb_prob_l"), t("mb_prob_u"),
                                                 t("mb_low"), t("mb_high"), mb, ptrs, [30] * C, t("mb_rep_t"))
    assert list(keys) == list(z["mb_keys"])
    assert [b[0].shape[0] for b in mb] == list(z["mb_bank_sizes"])
    assert abs(float(loss) - float(z["mb_loss"])) < 2e-5 * abs(float(z["mb_loss"]))
    loss.backward()
    assert np.abs(rep.grad.cpu().numpy() - z["mb_grad"]).max() <= 1e-4 * np.abs(z["mb_grad"]).max()
    mb, ptrs = banks()
    torch.manual_seed(78)
    proto, keys, loss = LH.compute_contra_memobank_loss(t("mb_rep"), t("mb_label_l"), t("mb_label_u"), t("mb_prob_l"),
                                                        t("mb_prob_u"), t("mb_low"), t("mb_high"), mb, ptrs, [30] * C,
                                                        t("mb_rep_t"), momentum_prototype=torch.ones(C, 256, 1, D, device=dev) * 0.1,
                                                        i_iter=5)
    assert abs(float(loss) - float(z["mb_loss_mom"])) < 2e-5 * abs(float(z["mb_loss_mom"]))
    assert np.abs(proto.sum((1, 2, 3)).cpu().numpy() - z["mb_proto_sum"]).max() < 1e-2


# ------------------------------------------------------------------ f4: ablation entry points (trian_CCT / trian_CPS)
def test_softmax_js_loss_matches_reference_function(dev, golden_dir):
    """cmlpl_softmax_js_f32 behind trian_CCT.softmax_js_loss against the reference's own function (trian_CCT.py:76-84,
    compiled from its unmodified source by oracle/make_golden.py), value and gradient, incl. exact-zero targets."""
    from cmlpl_b200.trian_CCT import cct_consistency, softmax_js_loss
    z = np.load(os.path.join(golden_dir, "ablation.npz"))
    for tag in ("a", "b"):
        x = torch.from_numpy(z[f"js_{tag}_z"]).to(dev).requires_grad_(True)
        t = torch.from_numpy(z[f"js_{tag}_t"]).to(dev)
        loss = softmax_js_loss(x, t)
        loss.backward()
        assert abs(float(loss) - float(z[f"js_{tag}_loss"])) <= 1e-5 * abs(float(z[f"js_{tag}_loss"]))
        g = z[f"js_{tag}_grad"]
        assert np.abs(x.grad.cpu().numpy() - g).max() <= 1e-4 * np.abs(g).max()
    with pytest.raises(AssertionError):
        softmax_js_loss(x.detach(), t)                       # the reference asserts inputs.requires_grad
    a = torch.randn(8, 9, device=dev, requires_grad=True); b = torch.randn(8, 9, device=dev, requires_grad=True)
    c = torch.randn(8, 9, device=dev, requires_grad=True)
    cct_consistency(a, b, c).backward()
    assert all(torch.isfinite(v.grad).all() for v in (a, b, c))


def test_cps_losses_match_reference(dev, golden_dir):
    from cmlpl_b200.trian_CPS import Distribution_Loss, cps_losses
    z = np.load(os.path.join(golden_dir, "ablation.npz"))
    ob = torch.from_numpy(z["cps_ob"]).to(dev).requires_grad_(True)
    oe = torch.from_numpy(z["cps_oe"]).to(dev).requires_grad_(True)
    tot, tot1, (cls, cls1, con, con1) = cps_losses(ob, oe, torch.from_numpy(z["cps_Y"]).to(dev))
    tot.backward(); tot1.backward()
    for got, want in ((tot, "cps_total"), (tot1, "cps_total1"), (con, "cps_con"), (con1, "cps_con1")):
        assert abs(float(got) - float(z[want])) <= 1e-5 * abs(float(z[want])), want
    assert np.abs(ob.grad.cpu().numpy() - z["cps_gb"]).max() <= 1e-5 * np.abs(z["cps_gb"]).max()
    assert np.abs(oe.grad.cpu().numpy() - z["cps_ge"]).max() <= 1e-5 * np.abs(z["cps_ge"]).max()
    with pytest.raises(NotImplementedError):
        Distribution_Loss(loss='mmd')(ob, oe)
