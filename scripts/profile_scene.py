"""A handful of full-scene inferences on the PaviaU-shaped synthetic scene, nothing else: the command
wrapped by `ncu` for the launch list and the `--set full` captures (GPU box only)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cmlpl_b200 import _lib, ops  # noqa: E402
from cmlpl_b200.tools.models import BaseNet2  # noqa: E402

_lib.require_device()
dev = torch.device("cuda")
R, C, B, K, w = 610, 340, 103, 9, 20
if len(sys.argv) > 5:
    R, C, B, K = (int(v) for v in sys.argv[2:6])
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 4
rng = np.random.default_rng(1088)
cube = torch.from_numpy(rng.standard_normal((R, C, 60)).astype(np.float32)).to(dev)
spectra = torch.from_numpy(rng.standard_normal((R * C, B)).astype(np.float32)).to(dev)
torch.manual_seed(1088)
net = BaseNet2(num_features=B, dropout=0, num_classes=K).to(dev).eval()
packed = net.packed_weights(w)
ws = ops.scene_workspace(R, C, B, K, w, dev)
labels = torch.empty(R * C, dtype=torch.uint8, device=dev)
torch.cuda.synchronize()
for _ in range(iters):
    ops.scene_infer(cube, spectra, packed, K, w, workspace=ws, labels=labels)
torch.cuda.synchronize()
print("ok", int(labels.sum()))
