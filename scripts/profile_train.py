"""A few fused mutual-learning steps (cmlpl_train_step, eager launches, PaviaU shape), nothing else: the command wrapped
by `ncu` for the launch list / metric captures of the training kernels (GPU box only)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cmlpl_b200 import _lib  # noqa: E402
from cmlpl_b200.fused_step import FusedMutualStep  # noqa: E402
from cmlpl_b200.tools.models import BaseNet2  # noqa: E402

_lib.require_device()
dev = torch.device("cuda")
B, K = 103, 9
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 3
torch.manual_seed(1088)
nets = [BaseNet2(B, 0.8, K).to(dev) for _ in range(2)]
g = torch.Generator(device=dev).manual_seed(1)
cube = torch.randn(610, 340, 60, device=dev, generator=g)
spectra = torch.randn(610 * 340, B, device=dev, generator=g)
pix = torch.randint(0, 610 * 340, (256,), device=dev, generator=g)
labels = torch.randint(0, K, (128,), device=dev, generator=g)
fs = FusedMutualStep(nets[0], nets[1], use_graph=False, thr=0.3)
torch.cuda.synchronize()
for it in range(iters):
    fs.queue_ptr, fs.queue_ptr1 = 0, 256
    fs.step(labels, 1, it, cube=cube, pix=pix, spectra=spectra)
torch.cuda.synchronize()
print("ok", [round(float(v), 4) for v in fs.hist[:5].cpu()])
