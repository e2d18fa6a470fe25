"""Two small scenes (w = 20 and w = 11) through cmlpl_scene_infer, meant to be run under
    compute-sanitizer --tool memcheck --error-exitcode 7 python scripts/sanitize_scene.py
(GPU box only; round 2: 0 errors for both window sizes)."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cmlpl_b200 import _lib, ops
from cmlpl_b200.tools.models import BaseNet2
_lib.require_device()
dev = torch.device('cuda')
for (R, C, B, K, w) in [(37, 45, 103, 9, 20), (23, 31, 224, 16, 11)]:
    rng = np.random.default_rng(1)
    cube = torch.from_numpy(rng.standard_normal((R, C, 60)).astype(np.float32)).to(dev)
    spectra = torch.from_numpy(rng.standard_normal((R * C, B)).astype(np.float32)).to(dev)
    torch.manual_seed(1)
    net = BaseNet2(B, 0, K, w=w).to(dev).eval()
    packed = net.packed_weights(w)
    lab = ops.scene_infer(cube, spectra, packed, K, w)
    torch.cuda.synchronize()
    print('ok', R, C, w, int(lab.sum()))
