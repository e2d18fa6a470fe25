"""Time patch_cnn alone (row-major and tiled P2) on a PaviaU-sized band (GPU box only)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from cmlpl_b200 import _lib
from cmlpl_b200.tools.models import BaseNet2
_lib.require_device()
R, C, w = 610, 340, 20
torch.manual_seed(0)
net = BaseNet2(103, 0, 9).cuda()
packed = net.packed_weights(w)
f0 = (torch.randn(8, R + w - 1, C + w - 1, 8, device="cuda") * 0.5).half()
n = R * C
p2 = torch.empty(((n + 127) // 128) * 128, 25, 64, dtype=torch.float16, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for name in ("cmlpl_patch_cnn_f16", "cmlpl_patch_cnn_f16_tiled"):
    for _ in range(3):
        _lib.call(name, f0.data_ptr(), C, w, R, packed.data_ptr(), p2.data_ptr(), st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        _lib.call(name, f0.data_ptr(), C, w, R, packed.data_ptr(), p2.data_ptr(), st)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{name}: {ms:.3f} ms  -> {n * 36.864e6 / ms / 1e9:.1f} TFLOP/s algorithmic, {n / ms / 1e3:.2f} Mpx/s")
