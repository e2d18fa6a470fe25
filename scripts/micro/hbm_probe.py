import torch
def t(fn, n=10):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
N = 16384 * 96000 // 4
a = torch.empty(N, device="cuda"); b = torch.randn(N, device="cuda")
ms = t(lambda: a.fill_(1.0)); print("fill  %.3f ms  %.0f GB/s written" % (ms, N * 4 / ms / 1e6))
ms = t(lambda: a.copy_(b)); print("copy  %.3f ms  %.0f GB/s read+write" % (ms, 2 * N * 4 / ms / 1e6))
ms = t(lambda: b.sum()); print("sum   %.3f ms  %.0f GB/s read" % (ms, N * 4 / ms / 1e6))
