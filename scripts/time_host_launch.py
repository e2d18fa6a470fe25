"""Host enqueue time vs device time of one scene inference, and the same step replayed from a CUDA graph (GPU box only)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from cmlpl_b200 import _lib, ops
from cmlpl_b200.tools.models import BaseNet2
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
step = lambda: ops.scene_infer(cube, spectra, packed, K, w, workspace=ws, labels=labels)
for _ in range(5): step()
torch.cuda.synchronize()
N = 50
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record()
for _ in range(N): step()
t1 = time.perf_counter(); e1.record(); torch.cuda.synchronize()
print(f"eager: host enqueue {1e3 * (t1 - t0) / N:.3f} ms/step, device {e0.elapsed_time(e1) / N:.3f} ms/step")
ref = labels.clone()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    step(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        step()
    for _ in range(5): g.replay()
    torch.cuda.synchronize()
    e0.record(s)
    for _ in range(N): g.replay()
    e1.record(s); torch.cuda.synchronize()
print(f"graph replay: device {e0.elapsed_time(e1) / N:.3f} ms/step, labels equal {bool(torch.equal(ref, labels))}")
