"""BASELINE configs[2] "contrastive batch 1024 (InfoNCE sim-matrix stress)": the loss kernels of the mutual-learning step
at btu = 1024 unlabelled rows, 1024-d unit features, 16 classes, T = 0.3 (f_s, f_w = normalize(randn[1024, 1024]),
z = randn[1024, 16], seeds 1088 / 1089).  Reports ms per call (CUDA events) and the HBM bytes each formulation moves for
  * cmlpl_graph_contrast_f32   (fp32: similarity GEMM -> [n, n] work buffer -> row kernel -> gradient GEMM)
  * cmlpl_bank_smooth_f32      (fp32: [n, queue] logits in a work buffer -> row kernel)
  * cmlpl_ntxent_f32           (NT-Xent at bs = 1024: [2bs, 2bs])
at n = 128 (the reference's batch) and n = 1024.  Called from bench.py --config c3."""
import torch
import torch.nn.functional as F


def run(dev, C=16):
    from cmlpl_b200 import ops
    out = {}
    for n in (128, 1024):
        g = torch.Generator().manual_seed(1088)
        fs = F.normalize(torch.randn(n, 1024, generator=g), dim=1).to(dev)
        g = torch.Generator().manual_seed(1089)
        fw = F.normalize(torch.randn(n, 1024, generator=g), dim=1).to(dev)
        z = torch.randn(n, C, generator=g).to(dev)
        p = torch.softmax(z, 1).contiguous(); p1 = torch.softmax(z.flip(0), 1).contiguous()
        queue = 10 * n
        qf = F.normalize(torch.randn(queue, 1024, generator=g).abs(), dim=1).to(dev)
        qp = torch.softmax(torch.randn(queue, C, generator=g), 1).to(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

        def t(fn, reps=20):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / reps

        r = {}
        from cmlpl_b200 import _lib
        zz = torch.cat([fs, fw]).contiguous()
        _lib.call("cmlpl_set_loss_gemm_mode", 1)            # similarity GEMMs on tcgen05 (fp16 operands)
        ref_l, _ = ops.graph_contrast(fs, fw, p1, p, 0.3, 0, 1.0, True)
        r["graph_contrast_tc_ms"] = t(lambda: ops.graph_contrast(fs, fw, p1, p, 0.3, 0, 1.0, True))
        r["bank_smooth_tc_ms"] = t(lambda: ops.bank_smooth(z, fw, qf, qp, 0.95, 0.3, True, 0.5))
        r["ntxent_tc_ms"] = t(lambda: ops.ntxent(zz, n, 0.5, True))
        _lib.call("cmlpl_set_loss_gemm_mode", 0)
        l32, _ = ops.graph_contrast(fs, fw, p1, p, 0.3, 0, 1.0, True)
        r["graph_contrast_tc_vs_fp32_rel"] = float((ref_l - l32).abs() / l32.abs())
        r["graph_contrast_ms"] = t(lambda: ops.graph_contrast(fs, fw, p1, p, 0.3, 0, 1.0, True))
        r["graph_contrast_matrix_bytes"] = 3 * n * n * 4
        r["graph_contrast_operand_bytes"] = 3 * n * 1024 * 4
        r["bank_smooth_ms"] = t(lambda: ops.bank_smooth(z, fw, qf, qp, 0.95, 0.3, True, 0.5))
        r["bank_smooth_matrix_bytes"] = n * queue * 4
        r["ntxent_ms"] = t(lambda: ops.ntxent(zz, n, 0.5, True))
        r["ntxent_matrix_bytes"] = 2 * (2 * n) ** 2 * 4
        r["note"] = ("matrices are produced by the GEMM and consumed by the row kernel that follows; at n = 1024 they are "
                     "4-16 MB, far below the 126 MB L2")
        out[f"n{n}"] = r
    return out
