"""Shared-conv1 path (conv1_scene -> patch_conv2) vs the per-patch kernel on the same conv0 map (GPU box)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from cmlpl_b200 import _lib, ops
from oracle import cmlpl_oracle as O
_lib.require_device()
dev = torch.device("cuda")
torch.manual_seed(3)
for (R, C) in ((37, 45), (610, 340)):
    w = 20; PR, PC = R + w - 1, C + w - 1; n = R * C
    sd = O.basenet2_init(103, 9)
    packed = ops.pack_basenet2({k: v.to(dev) for k, v in sd.items()}, 103, 9, w)
    f0 = (torch.randn(8, PR, PC, 8, device=dev) * 0.7).half()
    g = torch.empty(9, PR * PC, 64, device=dev)
    pm = torch.zeros(9, PR, PC, 64, dtype=torch.float16, device=dev)
    mt = (n + 127) // 128
    p2a = torch.zeros(mt, 200, 128, 8, dtype=torch.float16, device=dev)
    p2b = torch.zeros(mt, 200, 128, 8, dtype=torch.float16, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    def shared():
        _lib.call("cmlpl_conv1_scene_f16", f0.data_ptr(), C, w, R, packed.data_ptr(), g.data_ptr(), pm.data_ptr(), st)
        _lib.call("cmlpl_patch_conv2_f16_tiled", pm.data_ptr(), C, w, R, packed.data_ptr(), p2b.data_ptr(), st)
    _lib.call("cmlpl_patch_cnn_f16_tiled", f0.data_ptr(), C, w, R, packed.data_ptr(), p2a.data_ptr(), st)
    shared()
    torch.cuda.synchronize()
    a = p2a.float().permute(0, 2, 1, 3).reshape(mt * 128, 1600)[:n]
    b = p2b.float().permute(0, 2, 1, 3).reshape(mt * 128, 1600)[:n]
    d = (a - b).abs()
    print(f"{R}x{C}: P2 shared-conv1 vs per-patch: max|d| {float(d.max()):.3e} rel {float(d.max() / a.abs().max()):.2e}  exact-equal frac {float((a == b).float().mean()):.4f}")
    for _ in range(2): shared()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    torch.cuda.synchronize(); e0.record()
    for _ in range(5):
        _lib.call("cmlpl_conv1_scene_f16", f0.data_ptr(), C, w, R, packed.data_ptr(), g.data_ptr(), pm.data_ptr(), st)
    e1.record()
    for _ in range(5):
        _lib.call("cmlpl_patch_conv2_f16_tiled", pm.data_ptr(), C, w, R, packed.data_ptr(), p2b.data_ptr(), st)
    e2.record(); torch.cuda.synchronize()
    print(f"   conv1_scene+pool {e0.elapsed_time(e1) / 5:.3f} ms, patch_conv2 {e1.elapsed_time(e2) / 5:.3f} ms")
