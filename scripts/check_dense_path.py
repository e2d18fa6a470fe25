"""Dense conv2 / pool / classifier path, stage by stage, against torch on the same fp16 inputs and against
the per-pixel path (GPU box only; torch is the checker here, not the product)."""
import os, sys
import numpy as np, torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from cmlpl_b200 import _lib, ops
from oracle import cmlpl_oracle as O
import ctypes
_lib.require_device()
dev = torch.device("cuda")
REP = [0, 1, 4, 8, 9]
cls = lambda i: 0 if i == 0 else (2 if i == 9 else 1)


def rel(a, b):
    a = a.float(); b = b.float()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def run(R, C, B, K, seed, time_it=False):
    w = 20
    res = {}
    rng = np.random.default_rng(seed)
    cube = torch.from_numpy(rng.standard_normal((R, C, 60)).astype(np.float32)).to(dev)
    spectra = torch.from_numpy(rng.standard_normal((R * C, B)).astype(np.float32)).to(dev)
    torch.manual_seed(seed)
    sd = O.basenet2_init(B, K)
    sdd = {k: v.to(dev) for k, v in sd.items()}
    packed = ops.pack_basenet2(sdd, B, K, w)
    st = torch.cuda.current_stream().cuda_stream
    PR, PC = R + w - 1, C + w - 1
    PR2, PC2 = (R + w) // 2, (C + w) // 2
    n = R * C
    f0 = torch.empty(8 * PR * PC * 8, dtype=torch.float16, device=dev)
    g = torch.empty(9 * PR * PC * 64, dtype=torch.float32, device=dev)
    pm = torch.zeros(9 * PR * PC * 64, dtype=torch.float16, device=dev)
    pmq = torch.full((9, 4, 8, PR2, PC2, 8), float("nan"), dtype=torch.float16, device=dev)
    yq = torch.full((9, 4, 8, PR2, PC2, 8), float("nan"), dtype=torch.float16, device=dev)
    lmap = torch.full((4, 5, 4, PR2, PC2, 4), float("nan"), dtype=torch.float32, device=dev)
    _lib.call("cmlpl_conv0_map_f16", cube.data_ptr(), R, C, 0, R, w, 0, R, packed.data_ptr(), f0.data_ptr(), st)
    _lib.call("cmlpl_conv1_scene_f16", f0.data_ptr(), C, w, R, packed.data_ptr(), g.data_ptr(), pm.data_ptr(), st)
    _lib.call("cmlpl_conv1_scene_planes_f16", f0.data_ptr(), C, w, R, packed.data_ptr(), g.data_ptr(), pmq.data_ptr(), st)
    torch.cuda.synchronize()
    # ---- stage 1: planes == old pooled maps
    pm4 = pm.view(9, PR, PC, 64)
    pmq_ref = torch.zeros(9, 4, 8, PR2, PC2, 8, dtype=torch.float16, device=dev)
    for p in range(2):
        for q in range(2):
            sub = pm4[:, p:PR - 1:2, q:PC - 1:2, :]                    # pooled cells exist for pr < PR-1, pc < PC-1
            pmq_ref[:, p * 2 + q, :, :sub.shape[1], :sub.shape[2], :] = sub.reshape(9, sub.shape[1], sub.shape[2], 8, 8).permute(0, 3, 1, 2, 4)
    res["planes_equal"] = bool(torch.equal(pmq, pmq_ref))
    print(f"[{R}x{C}] planes vs pooled maps: equal={res['planes_equal']} nan={int(torch.isnan(pmq.float()).sum())}")
    # fused conv1+pool kernel (what scene_infer runs): same planes up to the fp32 summation order of the 2x2 pool
    pmq_f = torch.full((9, 4, 8, PR2, PC2, 8), float("nan"), dtype=torch.float16, device=dev)
    _lib.call("cmlpl_conv1_pool_planes_f16", f0.data_ptr(), C, w, R, packed.data_ptr(), pmq_f.data_ptr(), st)
    torch.cuda.synchronize()
    d = (pmq_f.float() - pmq.float()).abs()
    ulp = (pmq.float().abs() * 2.0 ** -10).clamp_min(2.0 ** -24)
    print(f"   fused conv1+pool planes: nan={int(torch.isnan(pmq_f.float()).sum())} bit-equal {float((pmq_f == pmq).float().mean()):.5f} "
          f"max diff in ulps {float((d / ulp).max()):.2f} max abs {float(d.max()):.3e}")
    pmq = pmq_f
    # dense planes [9][4][64][PR2][PC2] f32
    PMd = pmq_ref.permute(0, 1, 2, 5, 3, 4).reshape(9, 4, 64, PR2, PC2).float()
    # ---- stage 2: conv2 variants
    _lib.call("cmlpl_conv2_scene_f16", pmq.data_ptr(), C, w, R, packed.data_ptr(), yq.data_ptr(), st)
    torch.cuda.synchronize()
    W2 = sdd["conv2.weight"].half().float()
    b2 = sdd["conv2.bias"]
    Yd = torch.zeros(25, 4, 64, PR2, PC2, device=dev)
    for rho in range(5):
        for kap in range(5):
            acc = PMd[cls(REP[rho]) * 3 + cls(REP[kap])].clone() + b2.view(1, 64, 1, 1)
            for dy in range(3):
                i2 = REP[rho] + dy - 1
                if i2 < 0 or i2 > 9: continue
                for dx in range(3):
                    j2 = REP[kap] + dx - 1
                    if j2 < 0 or j2 > 9: continue
                    src = PMd[cls(i2) * 3 + cls(j2)]                                 # [4, 64, PR2, PC2]
                    sh = F.pad(src, (1, 1, 1, 1))[:, :, dy:dy + PR2, dx:dx + PC2]    # value at (y+dy-1, x+dx-1)
                    acc += torch.einsum("oc,pcyx->poyx", W2[:, :, dy, dx], sh)
            Yd[rho * 5 + kap] = acc.clamp_min(0)
    # the kernel stores 9 half-pooled maps: border classes averaged with their partner (rho 0 with rho 1 one row down, ...)
    ycl = lambda c, u: u if c == 0 else (2 if c == 1 else 3 + u)
    yq_d = yq.permute(0, 1, 2, 5, 3, 4).reshape(9, 4, 64, PR2, PC2).float()
    worst, nan_y = 0.0, 0
    for Al in range(3):
        for Be in range(3):
            us = [0, 1] if Al != 1 else [0]
            vs = [0, 1] if Be != 1 else [0]
            ref = torch.zeros(4, 64, PR2, PC2, device=dev)
            for u in us:
                for v in vs:
                    src = Yd[ycl(Al, u) * 5 + ycl(Be, v)]
                    ref += F.pad(src, (0, 1, 0, 1))[:, :, u:u + PR2, v:v + PC2] / (len(us) * len(vs))
            r0 = [0, 2, 3][Al]                                   # rows y' < rho of row class rho are never produced
            got = yq_d[Al * 3 + Be][:, :, r0:PR2 - 1, :PC2 - 1]
            nan_y += int(torch.isnan(got).sum())
            e = rel(torch.nan_to_num(got), ref[:, :, r0:PR2 - 1, :PC2 - 1])
            worst = max(worst, e)
            if not e < 5e-3:
                print("      map Al=%d Be=%d rel %.2e" % (Al, Be, e))
    res["half_pooled_rel"], res["half_pooled_nan"] = worst, nan_y
    print(f"   half-pooled conv2 maps: worst rel err {worst:.2e} (fp16 output rounding ~5e-4) nan={nan_y}")
    # ---- stage 3: pooled classifier partial maps from the kernel's own yq
    _lib.call("cmlpl_pool2_cls_f16", yq.data_ptr(), C, w, R, B, K, packed.data_ptr(), lmap.data_ptr(), st)
    torch.cuda.synchronize()
    Wc = sdd["classifier.weight"][:, :1600].reshape(K, 64, 5, 5).half().float()
    yz = torch.nan_to_num(yq_d)
    # the kernel sums the pooled columns: M[I][y, x] = sum_J L[I][J][y, x + 2J]
    worst = 0.0
    for I in range(5):
        Al = 0 if I == 0 else (2 if I == 4 else 1)
        ref = torch.zeros(4, K, PR2, PC2, device=dev)
        for J in range(5):
            Be = 0 if J == 0 else (2 if J == 4 else 1)
            us = [0, 1] if Al == 1 else [0]
            vs = [0, 1] if Be == 1 else [0]
            for u in us:
                for v in vs:
                    sh = F.pad(yz[Al * 3 + Be], (0, 2 * J + 1, 0, 1))[:, :, u:u + PR2, 2 * J + v:2 * J + v + PC2]
                    ref += torch.einsum("kc,pcyx->pkyx", Wc[:, :, I, J], sh) / (len(us) * len(vs))
        got = lmap[:, I].permute(0, 1, 4, 2, 3).reshape(4, 16, PR2, PC2)[:, :K]
        # map I is read at y' = r' + 2I >= 2I, x' = c' <= PC2 - 11 only; compare where every input exists
        e = rel(got[:, :, 2 * I:PR2 - 1, :PC2 - 10], ref[:, :, 2 * I:PR2 - 1, :PC2 - 10])
        worst = max(worst, e)
    res["row_maps_rel"] = worst
    print(f"   class-partial row maps: worst rel err {worst:.2e}")
    # ---- stage 4: whole path vs per-pixel path vs oracle
    labels, logits = ops.scene_infer(cube, spectra, packed, K, w, want_logits=True)
    torch.cuda.synchronize()
    # per-pixel path on the same conv0 map
    mt = (n + 127) // 128
    p2 = torch.empty(mt * 200 * 128 * 8, dtype=torch.float16, device=dev)
    kc = ((B + 15) // 16) * 2
    x16 = torch.empty(mt * kc * 1024, dtype=torch.float16, device=dev)
    h16 = torch.empty(mt * 128 * 1024, dtype=torch.float16, device=dev)
    lab2 = torch.empty(n, dtype=torch.uint8, device=dev); log2 = torch.empty(n, K, dtype=torch.float32, device=dev)
    _lib.call("cmlpl_patch_conv2_f16_tiled", pm.data_ptr(), C, w, R, packed.data_ptr(), p2.data_ptr(), st)
    _lib.call("cmlpl_spectral_hidden_tc", spectra.data_ptr(), n, B, K, w, packed.data_ptr(), x16.data_ptr(), h16.data_ptr(), st)
    _lib.call("cmlpl_head_tc", p2.data_ptr(), h16.data_ptr(), n, B, K, w, packed.data_ptr(), lab2.data_ptr(), log2.data_ptr(), st)
    torch.cuda.synchronize()
    res["dense_vs_pixel_rel"], res["labels_equal"] = rel(logits, log2), float((labels == lab2).float().mean())
    print(f"   logits dense vs per-pixel path: rel {res['dense_vs_pixel_rel']:.2e}, labels equal {res['labels_equal']:.5f}")
    if n <= 4000:
        lab_ref, log_ref = O.test_whole(sd, cube.cpu().numpy(), spectra.cpu().numpy(), w, return_logits=True)
        res["dense_vs_oracle_rel"] = rel(logits.cpu(), torch.from_numpy(log_ref))
        print(f"   logits dense vs oracle: rel {res['dense_vs_oracle_rel']:.2e}; per-pixel vs oracle {rel(log2.cpu(), torch.from_numpy(log_ref)):.2e}")
    if time_it:
        ws = ops.scene_workspace(R, C, B, K, w, dev)
        off = (ctypes.c_size_t * 12)()
        _lib.call("cmlpl_scene_workspace_layout", R, C, B, K, w, off)
        base = ws.data_ptr()
        names = ["conv0_map", "spectral_logits", "conv1_pool", "conv2_scene", "pool2_cls", "head_sum"]
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(7)]
        tot = np.zeros(6)
        for it in range(8):
            ev[5].record()
            _lib.call("cmlpl_conv0_map_f16", cube.data_ptr(), R, C, 0, R, w, 0, R, packed.data_ptr(), base + off[0], st)
            ev[6].record()
            _lib.call("cmlpl_spectral_logits_tc", spectra.data_ptr(), n, B, K, w, packed.data_ptr(), base + off[1], base + off[2], st)
            ev[0].record()
            _lib.call("cmlpl_conv1_pool_planes_f16", base + off[0], C, w, R, packed.data_ptr(), base + off[4], st)
            ev[1].record()
            _lib.call("cmlpl_conv2_scene_f16", base + off[4], C, w, R, packed.data_ptr(), base + off[5], st)
            ev[2].record()
            _lib.call("cmlpl_pool2_cls_f16", base + off[5], C, w, R, B, K, packed.data_ptr(), base + off[6], st)
            ev[3].record()
            _lib.call("cmlpl_head_sum_lmap", base + off[2], base + off[6], C, R, B, K, w, packed.data_ptr(), lab2.data_ptr(), None, st)
            ev[4].record()
            torch.cuda.synchronize()
            if it >= 3:
                tot += np.array([ev[5].elapsed_time(ev[6]), ev[6].elapsed_time(ev[0])] + [ev[i].elapsed_time(ev[i + 1]) for i in range(4)])
        print("   stage ms: " + ", ".join(f"{nm} {t / 5:.3f}" for nm, t in zip(names, tot)))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3): ops.scene_infer(cube, spectra, packed, K, w, workspace=ws, labels=lab2)
        e0.record()
        for _ in range(10): ops.scene_infer(cube, spectra, packed, K, w, workspace=ws, labels=lab2)
        e1.record(); torch.cuda.synchronize()
        print(f"   scene_infer {e0.elapsed_time(e1) / 10:.3f} ms  ({n / (e0.elapsed_time(e1) / 10) / 1e3:.1f} M px/s)")
    return res


if __name__ == "__main__":
    run(37, 45, 103, 9, 1)
    run(24, 75, 144, 15, 2)
    if "--big" in sys.argv:
        run(610, 340, 103, 9, 3, time_it=True)
