"""Where does the end-to-end time of StreamedRawScene go?  (GPU box only)"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from cmlpl_b200 import _lib, ops, preprocess, synth
from cmlpl_b200.tools.models import BaseNet2
from cmlpl_b200.tools.hyper_tools import StreamedRawScene
_lib.require_device()
dev = torch.device("cuda")
R, C, B, K, w = 610, 340, 103, 9, 20
cube_u16, _ = synth.synth_scene(R, C, B, K)
raw_host = torch.from_numpy(cube_u16.reshape(-1, B).copy()).pin_memory()
raw_dev = raw_host.to(dev)
pp = preprocess.fit(raw_dev, 60)
torch.manual_seed(0)
net = BaseNet2(B, 0, K).to(dev).eval(); packed = net.packed_weights(w)
cube, spec = preprocess.apply(raw_dev, pp)
ws = ops.scene_workspace(R, C, B, K, w, dev); labels = torch.empty(R * C, dtype=torch.uint8, device=dev)
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, (time.perf_counter() - w0) * 1e3 / n
print("scene_infer (device-resident)      %.3f ms gpu, %.3f ms wall" % t(lambda: ops.scene_infer(cube.view(R, C, 60), spec, packed, K, w, workspace=ws, labels=labels)))
print("preprocess.apply                   %.3f ms gpu, %.3f ms wall" % t(lambda: preprocess.apply(raw_dev, pp, cube=cube, spectra=spec)))
print("H2D raw (pinned -> device)         %.3f ms gpu, %.3f ms wall" % t(lambda: raw_dev.copy_(raw_host, non_blocking=True)))
lab_host = torch.empty(R * C, dtype=torch.uint8).pin_memory()
print("D2H labels                         %.3f ms gpu, %.3f ms wall" % t(lambda: lab_host.copy_(labels, non_blocking=True)))
for ns in (1, 2):
    st = StreamedRawScene(pp, R, C, B, K, w, nsplit=ns)
    print("StreamedRawScene nsplit=%d          %.3f ms gpu, %.3f ms wall" % ((ns,) + t(lambda: st(packed, raw_host))))
    print("StreamedRawScene nsplit=%d no d2h   %.3f ms gpu, %.3f ms wall" % ((ns,) + t(lambda: st(packed, raw_host, d2h=False))))
