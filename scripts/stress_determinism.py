"""Race hunt: the same scene inferred many times must give bit-identical logits every time (the dense kernels have no
atomics; any difference is a missing fence / barrier-phase bug).  GPU box only."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from cmlpl_b200 import _lib, ops
from oracle import cmlpl_oracle as O
_lib.require_device()
dev = torch.device("cuda")
bad = 0
for (R, C, B, K, reps) in [(610, 340, 103, 9, 300), (333, 517, 200, 16, 150), (97, 1031, 144, 15, 150), (211, 89, 224, 16, 200)]:
    gen = torch.Generator(device=dev); gen.manual_seed(R)
    cube = torch.randn((R, C, 60), device=dev, generator=gen)
    spectra = torch.randn((R * C, B), device=dev, generator=gen)
    torch.manual_seed(R)
    packed = ops.pack_basenet2({k: v.to(dev) for k, v in O.basenet2_init(B, K).items()}, B, K, 20)
    ws = ops.scene_workspace(R, C, B, K, 20, dev)
    lab0, log0 = ops.scene_infer(cube, spectra, packed, K, 20, want_logits=True, workspace=ws)
    lab0, log0 = lab0.clone(), log0.clone()
    diff = 0
    for i in range(reps):
        ws.random_(0, 255) if i % 7 == 0 else None            # garbage in the workspace must not matter
        lab, log = ops.scene_infer(cube, spectra, packed, K, 20, want_logits=True, workspace=ws)
        if not (torch.equal(lab, lab0) and torch.equal(log, log0)):
            diff += 1
    torch.cuda.synchronize()
    print(f"{R}x{C}x{B}/{K}: {reps} runs, {diff} differing")
    bad += diff
print("STRESS", "FAILED" if bad else "ok")
sys.exit(1 if bad else 0)
