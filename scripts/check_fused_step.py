"""Debug / numerics report of the fused training step (run on the GPU box):
every intermediate of cmlpl_train_step against torch fp32 on the same inputs, phase by phase.
    python scripts/check_fused_step.py [--bands 103 --classes 9]
"""
import argparse
import ctypes
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cmlpl_b200 import _lib, fused_step as FS  # noqa: E402
from cmlpl_b200.tools.models import BaseNet2  # noqa: E402


def rel(a, b):
    a = a.detach().double().cpu(); b = b.detach().double().cpu()
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-30))


def planes_to_nchw(buf, ns, npos, h):
    """[ns][8][npos][8] f16 -> [ns, 64, h, h] f32"""
    t = buf.view(torch.float16)[: ns * 8 * npos * 8].view(ns, 8, npos, 8).float()
    return t.permute(0, 1, 3, 2).reshape(ns, 64, h, h)


def bits_to_mask(buf, ns, npos, h):
    w = buf.view(torch.int32)[: ns * npos * 2].view(ns, npos, 2).cpu().numpy().astype(np.uint32)
    bits = ((w[..., None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(ns, npos, 64)
    return torch.from_numpy(bits.astype(np.float32)).permute(0, 2, 1).reshape(ns, 64, h, h)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bands", type=int, default=103)
    ap.add_argument("--classes", type=int, default=9)
    ap.add_argument("--bs", type=int, default=128)
    ap.add_argument("--dropout", type=float, default=0.8)
    ap.add_argument("--graph", action="store_true")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    _lib.require_device()
    torch.manual_seed(7)
    B, C, bs = a.bands, a.classes, a.bs
    nb = 2 * bs
    nets = [BaseNet2(B, a.dropout, C).to(dev) for _ in range(2)]
    g = torch.Generator().manual_seed(11)
    patches = torch.randn(2, nb, 60, 20, 20, generator=g).to(dev)
    spectra = torch.randn(2, nb, B, generator=g).to(dev)
    labels = torch.randint(0, C, (bs,), generator=g).to(dev)
    masks = (torch.rand(2, nb, 2624, generator=g) > a.dropout).float().to(dev) / (1 - a.dropout) if a.dropout > 0 else \
        torch.ones(2, nb, 2624, device=dev)
    fs = FS.FusedMutualStep(nets[0], nets[1], bs=bs, btu=bs, thr=0.15, num_epochs=20, use_graph=a.graph)
    for t in range(2):
        fs.queue_feats[t].copy_(F.normalize(torch.randn(fs.queue, 1024, generator=g).abs(), dim=1).to(dev))
        fs.queue_probs[t].copy_(torch.softmax(torch.randn(fs.queue, C, generator=g) * 2, 1).to(dev))
    qf = [q.clone() for q in fs.queue_feats]; qp = [q.clone() for q in fs.queue_probs]
    w_before = [[p.detach().clone() for p in ps] for ps in fs.params]
    fs.queue_ptr, fs.queue_ptr1 = 256, 512
    hist = fs.step(labels, epoch=1, batch_index=0, patches=patches, spectra=spectra, drop_masks=masks, phases=15)
    torch.cuda.synchronize()
    print("hist:", dict(zip(FS.HIST, [round(float(v), 5) for v in hist[:9].cpu()])))
    L = {}
    lay = (ctypes.c_size_t * 1)()
    # ---- torch reference on the same inputs (fp32, weights before the update)
    ref = []
    for e in range(2):
        sd = {k: w.clone().requires_grad_(True) for k, w in zip(FS.TENSORS, w_before[e])}
        x, y = patches[e], spectra[e]
        a0 = F.conv2d(x, sd["conv0.weight"], sd["conv0.bias"])
        a1 = F.relu(F.conv2d(a0, sd["conv1.weight"], sd["conv1.bias"], padding=1) + a0)
        p1 = F.avg_pool2d(a1, 2, 2)
        a2 = F.relu(F.conv2d(p1, sd["conv2.weight"], sd["conv2.bias"], padding=1) + p1)
        p2 = F.avg_pool2d(a2, 2, 2)
        h = F.relu(F.linear(y, sd["feat_spe.weight"], sd["feat_spe.bias"]))
        cat = torch.cat([p2.reshape(nb, -1), h], 1)
        feat = h / h.pow(2).sum(1, keepdim=True).sqrt()
        logits = F.linear(cat * masks[e], sd["classifier.weight"], sd["classifier.bias"])
        ref.append(dict(sd=sd, a0=a0, a1=a1, p1=p1, a2=a2, cat=cat, feat=feat, logits=logits))
    # workspace views
    lib = _lib.load()
    bsz = lambda n: n
    ws = fs.work
    lay = fs.workspace_layout()
    ns = 2 * nb
    x16 = planes_to_nchw(ws[lay["x16"]:], ns, 400, 20)
    print("x16 vs patches (60 ch):", rel(x16[:, :60], patches.view(ns, 60, 20, 20)), " pad ch max:", float(x16[:, 60:].abs().max()))
    a0 = planes_to_nchw(ws[lay["a0"]:], ns, 400, 20)
    a0r = torch.cat([r["a0"] for r in ref])
    print("a0:", rel(a0, a0r))
    p1 = planes_to_nchw(ws[lay["p1"]:], ns, 100, 10)
    print("p1:", rel(p1, torch.cat([r["p1"] for r in ref])))
    m1 = bits_to_mask(ws[lay["m1"]:], ns, 400, 20)
    a1r = torch.cat([r["a1"] for r in ref]).cpu()
    print("m1 disagreements:", int(((a1r > 0).float() != m1).sum()), "of", m1.numel())
    m2 = bits_to_mask(ws[lay["m2"]:], ns, 100, 10)
    a2r = torch.cat([r["a2"] for r in ref]).cpu()
    print("m2 disagreements:", int(((a2r > 0).float() != m2).sum()), "of", m2.numel())
    cat = ws[lay["cat"]:lay["cat"] + ns * 2624 * 4].view(torch.float32).view(ns, 2624)
    catr = torch.cat([r["cat"] for r in ref])
    print("cat conv part:", rel(cat[:, :1600], catr[:, :1600]), " spectral part:", rel(cat[:, 1600:], catr[:, 1600:]))
    print("logits:", rel(fs.logits.view(ns, C), torch.cat([r["logits"] for r in ref])),
          " feat:", rel(fs.feat.view(ns, 1024), torch.cat([r["feat"] for r in ref])))
    # ---- losses with torch (train.py:191-265)
    T, alpha, thr = fs.T, fs.alpha, fs.prm.adap_thr
    lo = [r["logits"] for r in ref]; fe = [r["feat"] for r in ref]
    with torch.no_grad():
        probs_o = [torch.softmax(lo[1][bs:], 1), torch.softmax(lo[0][bs:], 1)]
        fu = [fe[1][bs:], fe[0][bs:]]
        probs = []
        for t in range(2):
            A = torch.exp(fu[t] @ qf[t].t() / T); A = A / A.sum(1, keepdim=True)
            probs.append(alpha * probs_o[t] + (1 - alpha) * A @ qp[t])
        mask = [(p.max(1)[0] >= thr).float() for p in probs]
    print("probs:", rel(fs.probs[0], probs[0]), rel(fs.probs[1], probs[1]), " mask disagreements:",
          int((fs.mask[0] != mask[0]).sum() + (fs.mask[1] != mask[1]).sum()), " mask mean:", float(mask[0].mean()))
    cls = [F.cross_entropy(lo[e][:bs], labels) for e in range(2)]
    con = [(-(F.log_softmax(lo[e][bs:], 1) * probs[e]).sum(1) * mask[e]).mean() for e in range(2)]
    xs, xw = fe[0][bs:], fe[1][bs:]
    def contrast(fr, fc):
        sim = torch.exp(fr @ fc.t() / T); sp = sim / sim.sum(1, keepdim=True)
        Q0 = probs[1] @ probs[0].t(); Q0.fill_diagonal_(1)
        Q = Q0 * (Q0 >= 0.8).float(); Q = Q / Q.sum(1, keepdim=True)
        Qn = (1 - Q0) * (Q0 <= 0.3).float(); Qn = Qn / (Qn.sum(1, keepdim=True) + 1e-8)
        return (-(torch.log(sp) * Q).sum(1)).mean() + ((torch.log(sp + 1) * Qn).sum(1)).mean()
    lc = contrast(xs, xw.detach()); lc1 = contrast(xs.detach(), xw)
    tot = [cls[0] + 0.5 * lc + 4 * con[0], cls[1] + 0.5 * lc1 + 4 * con[1]]
    want = dict(lc=lc, total=tot[0], cls=cls[0], con=con[0], total1=tot[1], cls1=cls[1], con1=con[1], lc1=lc1)
    got = dict(zip(FS.HIST, hist[:9].cpu().tolist()))
    print("loss errors:", {k: abs(got[k] - float(v)) / max(abs(float(v)), 1e-9) for k, v in want.items()})
    tot[0].backward(); tot[1].backward()
    for e in range(2):
        errs = {k: rel(fs.grads[e][i], ref[e]["sd"][k].grad) for i, k in enumerate(FS.TENSORS)}
        print(f"grad errors net {e}:", {k: f"{v:.2e}" for k, v in errs.items()})
    # ---- linearised reference of the trunk backward: torch fp32 convolutions on the kernel's OWN saved tensors
    #      (dcat, ReLU masks, fp16 activations), so ReLU-mask flips of the fp16 forward do not enter the comparison
    prm = FS.TrainParams.from_buffer_copy(bytes(fs.prm_dev.cpu().numpy()))
    amax = prm.grad_amax
    import math as _m
    S = 2.0 ** (12 - _m.frexp(amax)[1]) if amax > 0 else 1.0
    print("grad_amax = %.4g  scale = 2^%d" % (amax, int(_m.log2(S))))
    dcat = ws[lay["dcat"]:lay["dcat"] + ns * 2624 * 4].view(torch.float32).view(ns, 2624)
    dp2 = dcat[:, :1600].reshape(ns, 64, 5, 5)
    m1d, m2d = m1.to(dev), m2.to(dev)
    dz2 = F.interpolate(dp2, scale_factor=2, mode="nearest") * 0.25 * m2d
    dz1_k = planes_to_nchw(ws[lay["dz1"]:], ns, 400, 20) / S
    da0_k = planes_to_nchw(ws[lay["da0"]:], ns, 400, 20) / S
    for e in range(2):
        sl = slice(e * nb, (e + 1) * nb)
        W2, W1 = w_before[e][4], w_before[e][2]
        dp1 = F.conv_transpose2d(dz2[sl], W2, padding=1) + dz2[sl]
        dz1 = F.interpolate(dp1, scale_factor=2, mode="nearest") * 0.25 * m1d[sl]
        da0 = F.conv_transpose2d(dz1, W1, padding=1) + dz1
        dW2 = torch.nn.grad.conv2d_weight(p1[sl], W2.shape, dz2[sl], padding=1)
        # each kernel is checked on the kernel-produced input of the previous one
        dz1i, da0i = dz1_k[sl], da0_k[sl]
        da0 = F.conv_transpose2d(dz1i, W1, padding=1) + dz1i
        dW1 = torch.nn.grad.conv2d_weight(a0[sl], W1.shape, dz1i, padding=1)
        dW0 = torch.nn.grad.conv2d_weight(x16[sl, :60], w_before[e][0].shape, da0i)
        print(f"linearised net {e}: dz1 {rel(dz1_k[sl], dz1):.2e} da0 {rel(da0_k[sl], da0):.2e} | dW2 {rel(fs.grads[e][4], dW2):.2e} "
              f"db2 {rel(fs.grads[e][5], dz2[sl].sum((0, 2, 3))):.2e} dW1 {rel(fs.grads[e][2], dW1):.2e} "
              f"db1 {rel(fs.grads[e][3], dz1i.sum((0, 2, 3))):.2e} dW0 {rel(fs.grads[e][0], dW0):.2e} "
              f"db0 {rel(fs.grads[e][1], da0i.sum((0, 2, 3))):.2e}")
        # per-sample view of the dz1 error
        err = (dz1_k[sl] - dz1).abs().amax((1, 2, 3)) / dz1.abs().amax()
        bad = (err > 1e-2).nonzero().flatten().tolist()
        print("   samples with dz1 error > 1e-2:", len(bad), bad[:20])
        if bad:
            b = bad[0]
            d = (dz1_k[sl][b] - dz1[b]).abs()
            print("   first bad sample: worst (c,y,x) =", np.unravel_index(int(d.argmax()), d.shape), "channels bad:",
                  int((d.amax((1, 2)) > 1e-2 * float(dz1.abs().amax())).sum()), "positions bad:",
                  int((d.amax(0) > 1e-2 * float(dz1.abs().amax())).sum()))
    # Adam: parameters after one step vs torch.optim.Adam on the reference gradients
    for e in range(2):
        ps = [ref[e]["sd"][k].detach().clone().requires_grad_(True) for k in FS.TENSORS]
        for p_, k in zip(ps, FS.TENSORS):
            p_.grad = ref[e]["sd"][k].grad.clone()
        torch.optim.Adam(ps, lr=5e-4).step()
        # first Adam step moves by lr*sign(g): compare where |g| is not tiny
        worst = 0.0
        for i, (p_, k) in enumerate(zip(ps, FS.TENSORS)):
            gz = ref[e]["sd"][k].grad.abs() > 1e-2 * ref[e]["sd"][k].grad.abs().max()
            worst = max(worst, float((fs.params[e][i].detach() - p_.detach())[gz].abs().max()))
        print(f"adam net {e}: max |dparam| where |g| > 1% max:", worst)
    # banks
    print("bank rows written:", float((fs.queue_feats[0][256:512] - torch.cat([fe[1][bs:], fe[0][:bs]]).detach()).abs().max()),
          float((fs.queue_feats[1][512:768] - torch.cat([fe[0][bs:], fe[1][:bs]]).detach()).abs().max()))
    # timing
    torch.cuda.synchronize()
    fs2 = FS.FusedMutualStep(nets[0], nets[1], bs=bs, btu=bs, thr=0.15, num_epochs=20, use_graph=True)
    for it in range(5):
        fs2.step(labels, 1, it, patches=patches, spectra=spectra, drop_masks=masks)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 200
    ev0.record()
    for it in range(n):
        list(fs2._graphs.values())[0][0].replay()
    ev1.record(); torch.cuda.synchronize()
    print("graph replay only: %.3f ms" % (ev0.elapsed_time(ev1) / n))
    fs3 = FS.FusedMutualStep(nets[0], nets[1], bs=bs, btu=bs, thr=0.15, num_epochs=20, use_graph=False)
    for ph in (1, 2, 4, 8, 15):
        for it in range(3):
            fs3.queue_ptr, fs3.queue_ptr1 = 0, 256
            fs3.step(labels, 1, it, patches=patches, spectra=spectra, drop_masks=masks, phases=ph)
        io = fs3._io(None, patches, spectra, None, None, None, masks)
        st = torch.cuda.current_stream().cuda_stream
        ev0.record()
        for it in range(50):
            _lib.call("cmlpl_train_step", ctypes.byref(io), ph, ctypes.c_void_p(st))
        ev1.record(); torch.cuda.synchronize()
        print("phases=%d (no graph, C call only): %.3f ms" % (ph, ev0.elapsed_time(ev1) / 50))
    ev0.record()
    for it in range(n):
        fs2.queue_ptr, fs2.queue_ptr1 = 0, 256
        fs2.step(labels, 1, it, patches=patches, spectra=spectra, drop_masks=masks)
    ev1.record(); torch.cuda.synchronize()
    print("fused step (graph): %.3f ms, launches/step = %d" % (ev0.elapsed_time(ev1) / n, FS.FusedMutualStep.launches()))
    import cProfile, pstats, io as _io
    pr = cProfile.Profile()
    pr.enable()
    for it in range(100):
        fs2.queue_ptr, fs2.queue_ptr1 = 0, 256
        fs2.step(labels, 1, it, patches=patches, spectra=spectra, drop_masks=masks)
    torch.cuda.synchronize()
    pr.disable()
    sio = _io.StringIO()
    pstats.Stats(pr, stream=sio).sort_stats("cumulative").print_stats(14)
    print(sio.getvalue()[:3000])


if __name__ == "__main__":
    main()
