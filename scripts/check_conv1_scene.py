"""Check the scene-level conv1 variants + pooled maps against torch (GPU box only)."""
import os, sys
import numpy as np, torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from cmlpl_b200 import _lib, ops
from oracle import cmlpl_oracle as O
_lib.require_device()
dev = torch.device("cuda")
torch.manual_seed(3)
R, C, w = 37, 45, 20
PR, PC = R + w - 1, C + w - 1
sd = O.basenet2_init(103, 9)
packed = ops.pack_basenet2({k: v.to(dev) for k, v in sd.items()}, 103, 9, w)
f0 = (torch.randn(8, PR, PC, 8, device=dev) * 0.7).half()
g = torch.zeros(9, PR * PC, 64, device=dev)
pm = torch.zeros(9, PR, PC, 64, dtype=torch.float16, device=dev)
st = torch.cuda.current_stream().cuda_stream
_lib.call("cmlpl_conv1_scene_f16", f0.data_ptr(), C, w, R, packed.data_ptr(), g.data_ptr(), pm.data_ptr(), st)
torch.cuda.synchronize()
x = f0.float().cpu().permute(0, 3, 1, 2).reshape(1, 64, PR, PC)          # [1, 64, PR, PC], channel = chunk*8 + e
W = sd["conv1.weight"].half().float(); b = sd["conv1.bias"]
S = {0: (1, 2), 1: (0, 1, 2), 2: (0, 1)}
G = torch.zeros(3, 3, 64, PR, PC)
worst = 0
for a in range(3):
    for bb in range(3):
        Wm = torch.zeros_like(W)
        for dy in S[a]:
            for dx in S[bb]:
                Wm[:, :, dy, dx] = W[:, :, dy, dx]
        G[a, bb] = F.relu(F.conv2d(x, Wm, b, padding=1) + x)[0]
        got = g[a * 3 + bb].cpu().reshape(PR, PC, 64).permute(2, 0, 1)
        err = float((got - G[a, bb]).abs().max() / G[a, bb].abs().max())
        worst = max(worst, err)
print("conv1_scene G variants: worst rel err %.2e" % worst)
worst = 0
for A in range(3):
    a0, a1 = (0 if A == 0 else 1), (2 if A == 2 else 1)
    for B in range(3):
        b0, b1 = (0 if B == 0 else 1), (2 if B == 2 else 1)
        ref = 0.25 * (G[a0, b0][:, :-1, :-1] + G[a0, b1][:, :-1, 1:] + G[a1, b0][:, 1:, :-1] + G[a1, b1][:, 1:, 1:])
        got = pm[A * 3 + B].float().cpu().permute(2, 0, 1)[:, :-1, :-1]
        worst = max(worst, float((got - ref).abs().max() / ref.abs().max()))
print("pool1_scene PM variants: worst rel err %.2e" % worst)
