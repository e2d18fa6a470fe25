"""Full-scene inference on every scene shape BASELINE.json names (configs[0..4]) on ONE B200: pixels/s of the device-resident
step (CUDA events) and parity of sampled pixels against the CPU oracle (GPU box only; random-normal inputs generated
on the device, random-init BaseNet2 under seed 1088).  configs[4] (8192 x 8192 x 224) is walked in row bands so the
intermediates stay bounded; it is run with w = 20 -- the reference's BaseNet2 hard-wires the classifier to 2624 inputs
(tools/models.py:127), so its 11 x 11 patch cannot go through the reference either.
    python scripts/bench_configs.py [--big]      -> one JSON line per config (also appended to gpurun_out/configs.jsonl)"""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from cmlpl_b200 import _lib, ops
from oracle import cmlpl_oracle as O          # checker only
_lib.require_device()
dev = torch.device("cuda")
W = 20
CONFIGS = [("C1 PaviaU", 610, 340, 103, 9, None), ("C2 Indian Pines", 145, 145, 200, 16, None),
           ("C3 Salinas", 512, 217, 204, 16, None), ("C4 Houston-2013", 349, 1905, 144, 15, None)]
if "--big" in sys.argv:
    CONFIGS.append(("C5 AVIRIS-NG-scale (w=20)", 8192, 8192, 224, 16, 256))


def rel(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


out = open(os.path.join(ROOT, "gpurun_out", "configs.jsonl"), "a") if os.path.isdir(os.path.join(ROOT, "gpurun_out")) else None
for name, R, C, B, K, band in CONFIGS:
    torch.manual_seed(1088)
    sd = O.basenet2_init(B, K)
    packed = ops.pack_basenet2({k: v.to(dev) for k, v in sd.items()}, B, K, W)
    gen = torch.Generator(device=dev); gen.manual_seed(1088)
    cube = torch.randn((R, C, 60), device=dev, generator=gen)
    n = R * C
    labels = torch.empty(n, dtype=torch.uint8, device=dev)
    band = R if band is None else band
    bands = [(a, min(a + band, R)) for a in range(0, R, band)]
    ws = ops.scene_workspace(min(band, R), C, B, K, W, dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    samples = {}
    if len(bands) == 1:
        spectra = torch.randn((n, B), device=dev, generator=gen)
        step = lambda: ops.scene_infer(cube, spectra, packed, K, W, workspace=ws, labels=labels)
        t_end = time.perf_counter() + 0.15
        while time.perf_counter() < t_end:
            step(); torch.cuda.synchronize()
        reps = 10
        e0.record()
        for _ in range(reps): step()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        lab, logits = ops.scene_infer(cube, spectra, packed, K, W, want_logits=True)
        idx = torch.randperm(n, device=dev, generator=gen)[:48].sort().values
        samples = {"idx": idx.cpu().numpy(), "spec": spectra[idx].cpu().numpy(), "logits": logits[idx].cpu().numpy()}
    else:
        ms, got_idx, got_spec, got_log = 0.0, [], [], []
        for bi, (a, b) in enumerate(bands):                       # spectra of a band are generated just before its call
            spectra = torch.randn(((b - a) * C, B), device=dev, generator=gen)
            logit_buf = None
            e0.record()
            ops.scene_infer(cube, spectra, packed, K, W, band_row0=a, band_rows=b - a, scene_rows=R, workspace=ws,
                            labels=labels[a * C:b * C])
            e1.record(); torch.cuda.synchronize()
            if bi > 0: ms += e0.elapsed_time(e1)                  # band 0 is the warm-up
            if bi in (0, len(bands) // 2, len(bands) - 1):        # parity samples from the first / middle / last band
                lab, logits = ops.scene_infer(cube, spectra, packed, K, W, band_row0=a, band_rows=b - a, scene_rows=R,
                                              workspace=ws, want_logits=True)
                j = torch.randperm((b - a) * C, device=dev, generator=gen)[:16].sort().values
                got_idx.append((j + a * C).cpu().numpy()); got_spec.append(spectra[j].cpu().numpy()); got_log.append(logits[j].cpu().numpy())
            del spectra
        ms = ms * len(bands) / (len(bands) - 1)                   # scale the timed bands to the whole scene
        samples = {"idx": np.concatenate(got_idx), "spec": np.concatenate(got_spec), "logits": np.concatenate(got_log)}
    # parity of the sampled pixels: patch = rows/cols r-10..r+9 through the a1 mirror map (hyper_tools.py:35-55,
    # 226-243; same map as oracle.extract_patches_at, applied per pixel so the 16 GB cube of C5 never leaves the device),
    # then the oracle's BaseNet2 fp32 forward on the host
    def mirror(o, m):
        return np.where(o < 0, -o - 1, np.where(o >= m, 2 * m - 1 - o, o))
    errs = []
    for i, p in enumerate(samples["idx"]):
        r, c = divmod(int(p), C)
        ri = torch.from_numpy(mirror(np.arange(r - 10, r + 10), R)).to(dev)
        ci = torch.from_numpy(mirror(np.arange(c - 10, c + 10), C)).to(dev)
        XP = cube[ri][:, ci].permute(2, 0, 1).contiguous().cpu()[None]
        with torch.no_grad():
            ref, _ = O.basenet2_forward(sd, XP, torch.from_numpy(samples["spec"][i:i + 1]))
        errs.append(rel(samples["logits"][i], ref.numpy()[0]))
    line = {"config": name, "rows": R, "cols": C, "bands": B, "classes": K, "w": W, "pixels": n, "row_bands": len(bands),
            "ms_per_scene": ms, "pixels_per_s": n / (ms / 1e3), "sampled_pixels": len(errs), "max_rel_logit_err_vs_oracle": max(errs)}
    print(json.dumps(line), flush=True)
    if out: out.write(json.dumps(line) + "\n"); out.flush()
    del cube, ws, labels
    torch.cuda.empty_cache()
