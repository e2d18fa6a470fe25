"""Stage-by-stage check of the scene-inference kernels against the CPU oracle (GPU box only).
Prints one line per stage; used while developing kernels (`gpurun -- python scripts/gpu_check.py`)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cmlpl_b200 import _lib, ops  # noqa: E402
from oracle import cmlpl_oracle as O  # noqa: E402


def rel(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def main():
    _lib.require_device()
    dev = torch.device("cuda")
    torch.manual_seed(1088)
    R, C, B, K, w = 23, 27, 103, 9, 20
    rng = np.random.default_rng(5)
    cube = rng.standard_normal((R, C, 60)).astype(np.float32)
    spectra = rng.standard_normal((R * C, B)).astype(np.float32)
    sd = O.basenet2_init(B, K)
    sd_dev = {k: v.to(dev) for k, v in sd.items()}
    packed = ops.pack_basenet2(sd_dev, B, K, w)
    torch.cuda.synchronize()
    print("pack ok", packed.numel())

    # ---- stage 1: conv0 map
    cube_d = torch.from_numpy(cube).to(dev)
    f0 = torch.empty((8, R + w - 1, C + w - 1, 8), dtype=torch.float16, device=dev)
    _lib.call("cmlpl_conv0_map_f16", cube_d.data_ptr(), R, C, 0, R, w, 0, R, packed.data_ptr(), f0.data_ptr(),
              torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    pad = np.pad(cube, ((10, 9), (10, 9), (0, 0)), mode="symmetric")
    ref0 = torch.einsum("rcf,of->rco", torch.from_numpy(pad), sd["conv0.weight"][:, :, 0, 0]) + sd["conv0.bias"]
    f0n = f0.permute(1, 2, 0, 3).reshape(R + w - 1, C + w - 1, 64)
    print("conv0_map rel err", rel(f0n.float().cpu().numpy(), ref0.numpy()))

    # ---- stage 3: patch_cnn on the device's own f0 (fp16) vs torch on the same fp16 values
    n = R * C
    p2 = torch.zeros((n, 25, 64), dtype=torch.float16, device=dev)
    t0 = time.time()
    _lib.call("cmlpl_patch_cnn_f16", f0.data_ptr(), C, w, R, packed.data_ptr(), p2.data_ptr(),
              torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    print("patch_cnn ran in %.3f s" % (time.time() - t0))
    f0c = f0n.float().cpu()
    w1 = sd["conv1.weight"].half().float(); w2 = sd["conv2.weight"].half().float()
    import torch.nn.functional as F
    idx = [0, 1, C - 1, C, n // 2, n - 1, 5 * C + 7]
    errs = []
    for p in idx:
        r, c = divmod(p, C)
        x0 = f0c[r:r + w, c:c + w, :].permute(2, 0, 1)[None]
        x1 = F.avg_pool2d(F.relu(F.conv2d(x0, w1, sd["conv1.bias"], padding=1) + x0), 2, 2)
        x1h = x1.half().float()
        x2 = F.avg_pool2d(F.relu(F.conv2d(x1h, w2, sd["conv2.bias"], padding=1) + x1h), 2, 2)
        refp = x2[0].permute(1, 2, 0).reshape(25, 64)
        errs.append(rel(p2[p].float().cpu().numpy(), refp.numpy()))
    print("patch_cnn rel errs (pixels %s):" % idx, ["%.2e" % e for e in errs])

    # ---- full scene
    spectra_d = torch.from_numpy(spectra).to(dev)
    labels, logits = ops.scene_infer(cube_d, spectra_d, packed, K, w, want_logits=True)
    torch.cuda.synchronize()
    lab_ref, log_ref = O.test_whole(sd, cube, spectra, w, return_logits=True)
    print("scene logits rel err", rel(logits.cpu().numpy(), log_ref), "label agreement",
          float(np.mean(labels.cpu().numpy() == lab_ref)))


if __name__ == "__main__":
    main()
