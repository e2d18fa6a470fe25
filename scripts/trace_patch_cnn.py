"""Dump the per-patch protocol timeline of CTA 0 of patch_cnn (diagnostics, GPU box only)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from cmlpl_b200 import _lib, ops
from cmlpl_b200.tools.models import BaseNet2

_lib.require_device()
R, C, w = 200, 340, 20
torch.manual_seed(0)
net = BaseNet2(103, 0, 9).cuda()
packed = net.packed_weights(w)
f0 = (torch.randn(8, R + w - 1, C + w - 1, 8, device="cuda") * 0.5).half()
p2 = torch.empty(R * C, 25, 64, dtype=torch.float16, device="cuda")
trace = torch.zeros(64, 16, dtype=torch.int64, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for _ in range(2):
    _lib.call("cmlpl_debug_patch_cnn_trace", f0.data_ptr(), C, w, R, packed.data_ptr(), p2.data_ptr(), trace.data_ptr(), st)
torch.cuda.synchronize()
t = trace.cpu().numpy()
names = ["ld:A1empty", "ld:done", "mma:A1full", "mma:c1issued", "mma:A2full", "mma:c2issued", "ep:C1full0", "ep:E1h0done",
         "ep:C1full1", "ep:E1h1done", "ep:preC2wait", "ep:C2full", "ep:E2done", "ep:ld1", "ep:pool1", "ep:st1"]
t0 = t[8, 2]
print("patch", " ".join(f"{n:>12s}" for n in names))
for p in range(8, 14):
    print(f"{p:5d}", " ".join(f"{int(t[p, k] - t0):12d}" for k in range(16)))
per = (t[40, 2] - t[8, 2]) / 32
print("cycles per patch (MMA A1full to A1full):", per)
