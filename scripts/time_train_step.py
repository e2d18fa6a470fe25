"""Time one mutual-learning step (GPU box only) and print the per-kernel share via torch profiler-free events."""
import os, sys, argparse, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from cmlpl_b200 import _lib, ops, train as T
_lib.require_device()
dev = torch.device("cuda")
R, C, B, K, w = 610, 340, 103, 9, 20
cube = torch.randn(R, C, 60, device=dev); spec = torch.randn(R * C, B, device=dev)
args = argparse.Namespace(temperature=0.3, thr=1.0, num_epochs=20, queue_batch=17, alpha=0.95, lr=5e-4, labeled_batch_size=128, dropout=0.8, noise=0.5)
st = T.make_state(B, K, args, dev)
Y = torch.randint(0, K, (128,), device=dev)
def step(i):
    idx = torch.randint(0, R * C, (256,), device=dev)
    def batch():
        z = torch.randn((256, 60, w, w), device=dev)
        return ops.patch_gather(cube, w, idx=idx, noise=z, noise_scale=0.5), spec[idx] + torch.randn((256, B), device=dev) * 0.5
    xb, sb = batch(); xe, se = batch()
    T.mutual_step(st, xb, sb, xe, se, Y, 1, i, args)
for i in range(3): step(i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record()
for i in range(10): step(3 + i)
e1.record(); torch.cuda.synchronize()
print("train step: %.3f ms (GPU events), %.3f ms wall" % (e0.elapsed_time(e1) / 10, (time.perf_counter() - t0) * 100))
