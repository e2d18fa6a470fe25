"""Does cutting a PaviaU-size scene into row bands (intermediates of a band closer to the 126 MB L2) beat one launch
sequence over the whole scene?  Times ops.scene_infer over 1..6 bands (halo rows are recomputed per band)."""
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
rng = np.random.default_rng(1088)
cube = torch.from_numpy(rng.standard_normal((R, C, 60)).astype(np.float32)).to(dev)
spectra = torch.from_numpy(rng.standard_normal((R * C, B)).astype(np.float32)).to(dev)
torch.manual_seed(1088)
net = BaseNet2(num_features=B, dropout=0, num_classes=K).to(dev).eval()
packed = net.packed_weights(w)
labels = torch.empty(R * C, dtype=torch.uint8, device=dev)
ref = ops.scene_infer(cube, spectra, packed, K, w).clone()
for nb in (1, 2, 3, 4, 6):
    per = -(-R // nb)
    bands = [(a, min(a + per, R)) for a in range(0, R, per)]
    ws = ops.scene_workspace(per, C, B, K, w, dev)

    def step():
        for a, b in bands:
            ops.scene_infer(cube, spectra[a * C:b * C], packed, K, w, band_row0=a, band_rows=b - a, workspace=ws,
                            labels=labels[a * C:b * C])
    for _ in range(200):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(300):
        step()
    e1.record()
    torch.cuda.synchronize()
    print(f"{nb} bands of {per} rows: {e0.elapsed_time(e1) / 300:.4f} ms per scene, labels equal {bool(torch.equal(labels, ref))}")
