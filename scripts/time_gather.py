"""Time the materialising patch gather (random and contiguous pixels) against the HBM roofline."""
import os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from cmlpl_b200 import _lib, ops
_lib.require_device()
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
R, C, w = 610, 340, 20
cube = torch.randn(R, C, 60, device="cuda")
ng = 16384
out = torch.empty(ng, 60, w, w, device="cuda")
for name, idx in (("random", torch.randperm(R * C, device="cuda")[:ng].contiguous()), ("contiguous", torch.arange(50000, 50000 + ng, device="cuda"))):
    for _ in range(3): ops.patch_gather(cube, w, idx=idx, out=out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(10): ops.patch_gather(cube, w, idx=idx, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    gbs = ng * 96240 / ms / 1e6
    print(f"gather {name}: {ms:.3f} ms, {gbs:.0f} GB/s = {gbs / peak:.2%} of measured HBM peak, {ng / ms / 1e3:.1f} Mpx/s")
