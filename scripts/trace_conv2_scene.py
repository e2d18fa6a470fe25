"""Timeline of one conv2_scene_kernel CTA (clock64 stamps of the MMA issuer, the two epilogue groups and the loader).
Needs the instrumented library:  make -C cmlpl_b200/csrc trace   (-> scripts/_trace/libcmlpl_trace.so, git-ignored);
GPU box only, debugging aid -- nothing in the product loads that library."""
import ctypes
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
ws = ops.scene_workspace(R, C, B, K, w, dev)
labels = torch.empty(R * C, dtype=torch.uint8, device=dev)
ops.scene_infer(cube, spectra, packed, K, w, workspace=ws, labels=labels)
torch.cuda.synchronize()
off = (ctypes.c_size_t * 12)()
_lib.call("cmlpl_scene_workspace_layout", R, C, B, K, w, off)
tl = ctypes.CDLL(os.path.join(ROOT, "scripts", "_trace", "libcmlpl_trace.so"))
tl.cmlpl_conv2_scene_f16.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    rc = tl.cmlpl_conv2_scene_f16(ws.data_ptr() + off[4], C, w, R, packed.data_ptr(), ws.data_ptr() + off[5], st)
    assert rc == 0
torch.cuda.synchronize()
buf = np.zeros((4, 4096), dtype=np.uint64)
assert tl.cmlpl_debug_c2s_trace(buf.ctypes.data_as(ctypes.c_void_p)) == 0
t0 = int(min(buf[r, 0] for r in range(4) if buf[r, 0]))
names = ["mma", "epi0", "epi1", "load"]
ev = []
for r in range(4):
    for i in range(2047):
        c, tag = int(buf[r, 2 * i]), int(buf[r, 2 * i + 1])
        if c == 0:
            break
        ev.append((c - t0, names[r], tag))
ev.sort()
G = ["PA", "T1", "T2", "T3", "PB"]
tile = int(sys.argv[1]) if len(sys.argv) > 1 else 5
# the MMA warp issues 25 groups per tile: find the clock window of tile `tile`
mma = [(c, t) for c, n, t in ev if n == "mma"]
lo, hi = mma[tile * 50][0], mma[min((tile + 1) * 50 + 10, len(mma) - 1)][0]
print("cycles per tile (MMA issue start to start):", [mma[(i + 1) * 50][0] - mma[i * 50][0] for i in range(min(12, len(mma) // 50 - 1))])
for c, n, t in ev:
    if c < lo or c > hi:
        continue
    if n == "mma":
        d = f"kap{t >> 4} {G[(t & 15) >> 1]} {'issued' if t & 1 else 'start'}"
    elif n == "load":
        d = f"tma tile {t}"
    else:
        r = t >> 2
        d = f"item kap{r // 3} it{r % 3} " + ["wait", "ready", "done"][t & 3]
    print(f"{c - lo:7d} {n:5s} {d}")
