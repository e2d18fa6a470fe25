"""A handful of full-scene inferences through the RAW-cube entry point (cmlpl_scene_infer_raw, uint16 PaviaU-shaped
scene): the command wrapped by `ncu` for the per-kernel table of the end-to-end path (GPU box only)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cmlpl_b200 import _lib, ops, preprocess, synth  # noqa: E402
from cmlpl_b200.tools.models import BaseNet2  # noqa: E402

_lib.require_device()
dev = torch.device("cuda")
R, C, B, K, w = 610, 340, 103, 9, 20
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 4
cube, _ = synth.synth_scene(R, C, B, K, seed=1088)
raw = torch.from_numpy(cube.reshape(-1, B)).to(dev)
pp = preprocess.fit(raw, 60)
torch.manual_seed(1088)
net = BaseNet2(num_features=B, dropout=0, num_classes=K).to(dev).eval()
packed = net.packed_weights(w)
folded = pp.folded_conv0(net.conv0.weight, net.conv0.bias, dev)
ws = ops.scene_workspace(R, C, B, K, w, dev)
labels = torch.empty(R * C, dtype=torch.uint8, device=dev)
torch.cuda.synchronize()
for _ in range(iters):
    ops.scene_infer_raw(raw, folded, packed, K, C, w, workspace=ws, labels=labels)
torch.cuda.synchronize()
print("ok", int(labels.sum()))
