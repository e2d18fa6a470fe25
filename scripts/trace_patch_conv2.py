import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from cmlpl_b200 import _lib
from cmlpl_b200.tools.models import BaseNet2
_lib.require_device()
R, C, w = 200, 340, 20
PR, PC = R + w - 1, C + w - 1
net = BaseNet2(103, 0, 9).cuda(); packed = net.packed_weights(w)
pm = (torch.randn(9, PR, PC, 64, device="cuda") * 0.5).half()
n = R * C
p2 = torch.empty((n + 127) // 128, 200, 128, 8, dtype=torch.float16, device="cuda")
trace = torch.zeros(64, 16, dtype=torch.int64, device="cuda")
import sys as _s
mode = int(_s.argv[1]) if len(_s.argv) > 1 else 0
trace.view(-1)[1023] = mode
st = torch.cuda.current_stream().cuda_stream
for _ in range(2):
    _lib.call("cmlpl_debug_patch_conv2_trace", pm.data_ptr(), C, w, R, packed.data_ptr(), p2.data_ptr(), trace.data_ptr(), st)
torch.cuda.synchronize()
t = trace.cpu().numpy(); t0 = t[8, 2]
names = ["ld:Aempty", "ld:done", "mma:Afull", "mma:Dempty", "mma:issued", "ep:top", "ep:preD", "ep:Dfull", "ep:done"]
print("pair ", " ".join(f"{n:>11s}" for n in names))
for p in range(8, 14):
    print(f"{p:5d}", " ".join(f"{int(t[p, k] - t0):11d}" for k in range(9)))
print("mode", mode, "cycles per pair:", (t[40, 2] - t[8, 2]) / 32)
