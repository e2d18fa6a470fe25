#!/usr/bin/env python
"""Headline benchmark: pixels/sec of full-scene inference on a PaviaU-shaped synthetic scene
(610 x 340 x 103, 9 classes; BASELINE.json metric / configs[0]) on N B200s of one node.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle port)

One "step" = one pass of the hot path over the whole scene: conv0 map -> spectral branch ->
scene-level tcgen05 conv1+pool / conv2 (exact compute sharing) -> pool + classifier partial maps -> sum head + argmax (+ label-map
all-gather and confusion all-reduce when N > 1).  `value` is device-resident throughput; `e2e` goes through the public call with
pinned HOST buffers (H2D of the cube + spectra and D2H of the label map inside the timed region).
Multi-GPU: row bands, weak scaling -- the scene grows to (610*N) x 340 and every rank infers a
610-row band (+ read-only halo); no data-path collective.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

R0, C0, B0, K0, W0 = 610, 340, 103, 9, 20
FLOP_PER_PX_CONV2 = 2 * 3_686_400                       # conv2 MACs per pixel: the per-pixel kernel
FLOP_PER_PX_CONV1 = 2 * 14_745_600                      # conv1 MACs per pixel in the reference's per-patch arithmetic
FLOP_PER_PX_ALL = 2 * (1_536_000 + 14_745_600 + 3_686_400 + 1024 * B0 + 2624 * K0)   # 40.19 MFLOP (8d)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        z = json.load(open(p))
        return {"hbm": z["hbm_gbs"], "tf_burst": z["bf16_tflops"], "tf_sustained": z["bf16_tflops_sustained"],
                "src": "measured"}
    return {"hbm": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "src": "fallback"}


# ------------------------------------------------------------------ clocks sampler
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(self.gpu)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm = [float(r[1]) for r in rows if len(r) >= 8]
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in rows if len(r) >= 8 for i in range(4) if r[4 + i].strip() == "Active"})
        busy = [s for s in sm if s > 0.5 * max(sm)] or sm
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": float(rows[0][2]), "reasons": reasons,
                "samples": len(sm), "power_w_max": max(float(r[3]) for r in rows if len(r) >= 8)}


# ------------------------------------------------------------------ CPU arm (oracle port of the reference path)
def cpu_reference_pass(cube_pca, spectra, sd, rows, threads):
    """The reference's test_whole path on host cores for scene rows [rows[0], rows[1]):
    per-pixel window copy out of the mirror-padded cube (hyper_tools.py:231-243), batches of 512
    through BaseNet2 in fp32 (models.py:130-152), argmax (hyper_tools.py:426).  Returns seconds."""
    from oracle import cmlpl_oracle as O          # CPU baseline leg only

    torch.set_num_threads(threads)
    R, C, F = cube_pca.shape
    w, hw = W0, W0 // 2
    Xm = O.mirrow_cut(cube_pca, hw)               # one-off for the scene: not timed (amortised over all pixels)
    r0, r1 = rows
    n = (r1 - r0) * C
    t0 = time.perf_counter()
    XP = np.zeros((n, w, w, F), dtype=np.float32)
    k = 0
    for r in range(r0, r1):
        for c in range(C):
            XP[k] = Xm[r:r + 2 * hw, c:c + 2 * hw, :]
            k += 1
    XP = np.moveaxis(XP, 3, 1).astype(np.float32)
    Xs = spectra[r0 * C:r1 * C]
    out = []
    with torch.no_grad():
        for s in range(0, n, 512):
            xp = torch.from_numpy(XP[s:s + 512].astype("float32").copy())
            xs = torch.from_numpy(Xs[s:s + 512].astype("float32").copy())
            lo, _ = O.basenet2_forward(sd, xp, xs)
            out.append(torch.max(lo, 1)[1].numpy())
    np.concatenate(out)
    return time.perf_counter() - t0, n


def run_reference_arm(args, scene):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cmlpl_oracle as O
    cube_pca, spectra, _ = scene
    torch.manual_seed(1088)
    sd = O.basenet2_init(B0, K0)
    threads = os.cpu_count() or 1
    rows_per_step = 6
    times, npx = [], 0
    for i in range(args.warmup + args.steps):
        r0 = (37 + i * rows_per_step) % (R0 - rows_per_step)
        dt, npx = cpu_reference_pass(cube_pca, spectra, sd, (r0, r0 + rows_per_step), threads)
        if i >= args.warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    val = npx / (ms / 1e3)
    line = {
        "impl": "reference", "metric": "pixels/sec full-scene inference", "value": val, "unit": "pixels/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus, note="CPU arm: one step = a 6-row band (2040 px) of the same scene"),
        "cpu_baseline": {"value": val, "unit": "pixels/s", "cores": threads, "kind": "port",
                         "sample": f"{rows_per_step} rows x {C0} cols = {npx} px per step through the oracle port of "
                                   "ExtractPatches loop + BaseNet2 fp32 (bs 512) + argmax"},
        "e2e": {"value": val, "unit": "pixels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(n_gpus, note=None):
    cfg = {"workload": f"PaviaU-shaped synthetic scene {R0}x{C0}x{B0}, {K0} classes, w={W0}, n_PC=60 "
                       "(BASELINE.json configs[0], the configuration the metric is quoted on); "
                       "random-init BaseNet2 (seed 1088), synthetic cube seed 1088",
           "scene_rows": R0 * n_gpus, "scene_cols": C0, "bands": B0, "classes": K0, "patch": W0,
           "pixels_per_step": R0 * C0 * n_gpus,
           "parallelism": f"row bands x{n_gpus} (weak: {R0} rows + halo per GPU, scene = {R0 * n_gpus} rows)",
           "l2_policy": "per-step working set (cube 50 MB + spectra 85 MB + conv0 map 29 MB + spectral tiles 47 MB + pooled "
                        "planes 263 MB + conv2 variants 730 MB + class-partial maps 365 MB + partial spectral logits 53 MB) "
                        "exceeds the 126 MB L2; no explicit flush"}
    cfg["prewarm"] = "each timed loop is preceded by ~0.15 s of untimed steps (SM clock ramp from idle) and the W warm-up steps"
    if note:
        cfg["note"] = note
    return cfg


# ------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cmlpl_b200", choices=["cmlpl_b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup

    from cmlpl_b200 import synth
    t_data = time.time()
    scene4 = synth.preprocessed_scene(R0, C0, B0, K0, 60, 1088, return_raw=True)
    scene, raw_u16 = scene4[:3], scene4[3]
    t_data = time.time() - t_data
    if args.impl == "reference":
        return run_reference_arm(args, scene)

    import torch.distributed as dist
    from cmlpl_b200 import _lib, ops, parallel
    from cmlpl_b200.tools.models import BaseNet2

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    _lib.require_device()
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    cube_pca, spectra, gt = scene

    # ---- weak scaling: the scene is the base scene stacked `world` times; this rank's band = copy `rank`
    scene_rows = R0 * world
    r0, r1 = parallel.band_of(rank, world, scene_rows)
    s0, s1 = parallel.slab_of(r0, r1, scene_rows, W0)
    slab_host = torch.from_numpy(np.ascontiguousarray(cube_pca[np.arange(s0, s1) % R0])).pin_memory()
    spec_host = torch.from_numpy(np.ascontiguousarray(spectra)).pin_memory()      # band rows == base scene rows
    truth = torch.from_numpy(gt.reshape(-1).astype(np.int64) - 1).to(dev)       # -1 = unlabelled (ignored)
    n_band = (r1 - r0) * C0

    torch.manual_seed(1088)
    net = BaseNet2(num_features=B0, dropout=0, num_classes=K0).to(dev).eval()
    packed = net.packed_weights(W0)
    slab = slab_host.to(dev)
    spec = spec_host.to(dev)
    ws = ops.scene_workspace(r1 - r0, C0, B0, K0, W0, dev)
    labels = torch.empty(n_band, dtype=torch.uint8, device=dev)
    labels_host = torch.empty(n_band, dtype=torch.uint8).pin_memory()
    cm = torch.zeros(K0, K0, dtype=torch.int64, device=dev)

    def step_device():
        ops.scene_infer(slab, spec, packed, K0, W0, band_row0=r0, band_rows=r1 - r0, scene_rows=scene_rows,
                        slab_row0=s0, workspace=ws, labels=labels)
        if world > 1:
            parallel.gather_label_map(labels, scene_rows, C0)
            cm.zero_()
            parallel.reduce_confusion(ops.confusion(labels, truth, K0, cm))

    from cmlpl_b200.tools.hyper_tools import StreamedScene
    streamed = StreamedScene(scene_rows, C0, B0, K0, W0, nsplit=4, row0=r0, rows=r1 - r0, device=dev)
    assert (streamed.s0, streamed.s1) == (s0, s1)

    def step_e2e_f32():
        # host cube slab + spectra (already preprocessed, f32) in, uint8 labels out
        if world > 1:
            lab = streamed(packed, slab_host, spec_host, d2h=False)
            parallel.gather_label_map(lab, scene_rows, C0)
            streamed.labels_host.copy_(lab, non_blocking=True)
        else:
            streamed(packed, slab_host, spec_host)

    # headline end-to-end call: the RAW uint16 cube in pinned host memory in, uint8 labels out.  The
    # preprocessing parameters (band means/stds, PCA basis) are fitted once beforehand (on device);
    # per step: band-wise H2D of the raw rows overlapped with compute, z-score + PCA projection on
    # device, scene inference, D2H of the label map.
    from cmlpl_b200 import preprocess
    from cmlpl_b200.tools.hyper_tools import StreamedRawScene
    raw_rows = np.ascontiguousarray(raw_u16[np.arange(s0, s1) % R0].reshape(-1, B0))
    raw_host = torch.from_numpy(raw_rows).pin_memory()
    pp = preprocess.fit(torch.from_numpy(np.ascontiguousarray(raw_u16.reshape(-1, B0))).to(dev), 60)
    streamed_raw = StreamedRawScene(pp, scene_rows, C0, B0, K0, W0, nsplit=1, row0=r0, rows=r1 - r0, device=dev)
    folded = pp.folded_conv0(net.conv0.weight, net.conv0.bias, dev)

    def step_e2e():
        if world > 1:
            lab = streamed_raw(packed, raw_host, d2h=False, folded=folded)
            parallel.gather_label_map(lab, scene_rows, C0)
            streamed_raw.sets[0]["labels_host"].copy_(lab, non_blocking=True)
        else:
            streamed_raw(packed, raw_host, folded=folded)

    def timed(fn, steps, warmup):
        # a 20-step loop lasts ~20 ms: without load beforehand the SM clock is still ramping up from idle during it, so
        # every timed loop is preceded by ~0.15 s of untimed steps (recorded in config.prewarm), then the W warm-up steps
        t_end = time.perf_counter() + 0.15
        while time.perf_counter() < t_end:
            for _ in range(5):
                fn()
            torch.cuda.synchronize()
        for _ in range(warmup):
            fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps

    sampler = ClockSampler(local)
    sampler.start()
    ms_dev = timed(step_device, args.steps, args.warmup)
    ms_e2e = timed(step_e2e, args.steps, args.warmup)
    ms_e2e_f32 = timed(step_e2e_f32, args.steps, args.warmup)

    # ---- per-kernel durations of the same step, CUDA events on the launching stream
    import ctypes
    L = _lib.load()
    st = torch.cuda.current_stream().cuda_stream
    nb = r1 - r0
    off = (ctypes.c_size_t * 12)()          # f0pad, x16, h16, g, pmq, yq, lmap, p2, spe, hidden, total, tc
    _lib.call("cmlpl_scene_workspace_layout", nb, C0, B0, K0, W0, off)
    assert off[10] == ws.numel() and off[11] == 1, "bench expects the tensor-core scene path"
    o_f0, o_x16, o_h16, o_g, o_pmq, o_yq, o_lmap = (ws.data_ptr() + off[i] for i in range(7))
    pk = packed.data_ptr()
    names = ["conv0_map", "spectral_logits", "conv1_pool", "conv2_scene", "pool2_cls", "head_sum"]
    NS = len(names)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(NS + 1)] for _ in range(args.steps)]
    for it in range(args.warmup + args.steps):
        e = ev[it - args.warmup] if it >= args.warmup else [None] * (NS + 1)
        if e[0]: e[0].record()
        _lib.call("cmlpl_conv0_map_f16", slab.data_ptr(), scene_rows, C0, s0, s1 - s0, W0, r0, nb, pk, o_f0, st)
        if e[1]: e[1].record()
        _lib.call("cmlpl_spectral_logits_tc", spec.data_ptr(), n_band, B0, K0, W0, pk, o_x16, o_h16, st)
        if e[2]: e[2].record()
        _lib.call("cmlpl_conv1_pool_planes_f16", o_f0, C0, W0, nb, pk, o_pmq, st)
        if e[3]: e[3].record()
        _lib.call("cmlpl_conv2_scene_f16", o_pmq, C0, W0, nb, pk, o_yq, st)
        if e[4]: e[4].record()
        _lib.call("cmlpl_pool2_cls_f16", o_yq, C0, W0, nb, B0, K0, pk, o_lmap, st)
        if e[5]: e[5].record()
        _lib.call("cmlpl_head_sum_lmap", o_h16, o_lmap, C0, nb, B0, K0, W0, pk, labels.data_ptr(), None, st)
        if e[6]: e[6].record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    stage_ms = {names[i]: float(np.mean([e[i].elapsed_time(e[i + 1]) for e in ev])) for i in range(NS)}

    # ---- the HBM-bound kernel of the path: materialising patch gather (training batches / ExtractPatches)
    gather = None
    if world == 1:
        ng = 16384
        gout = torch.empty((ng, 60, W0, W0), dtype=torch.float32, device=dev)
        gres = {}
        for gname, gidx in (("contiguous", torch.arange(50000, 50000 + ng, device=dev)),
                            ("random", torch.randperm(n_band, device=dev)[:ng].contiguous())):
            for _ in range(3):
                ops.patch_gather(slab, W0, idx=gidx, out=gout)
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            g0.record()
            for _ in range(10):
                ops.patch_gather(slab, W0, idx=gidx, out=gout)
            g1.record()
            torch.cuda.synchronize()
            gres[gname] = g0.elapsed_time(g1) / 10
        gms = gres["contiguous"]
        gather = {"kernel": "patch_gather_reg_kernel (16384 raster-consecutive pixels = ExtractPatches over scene rows, "
                            "60x20x20 f32 each; output 1.57 GB > L2)",
                  "bound": "hbm", "ms": gms, "bytes_per_pixel": 96240, "ms_random_pixels": gres["random"]}
        del gout

    # ---- second half of the BASELINE metric: one mutual-learning train step (train.py:149-278), 128+128
    train_ms = None
    if world == 1:
        import argparse as _ap
        from cmlpl_b200 import train as T
        targs = _ap.Namespace(temperature=0.3, thr=1.0, num_epochs=20, queue_batch=17, alpha=0.95, lr=5e-4,
                              labeled_batch_size=128, dropout=0.8, noise=0.5)
        tst = T.make_state(B0, K0, targs, dev)
        lab_idx = torch.nonzero(truth >= 0).flatten()
        def train_step(i):
            li = lab_idx[torch.randint(0, lab_idx.numel(), (128,), device=dev)]
            ui = lab_idx[torch.randint(0, lab_idx.numel(), (128,), device=dev)]
            both = torch.cat([li, ui])
            def batch():
                z = torch.randn((256, 60, W0, W0), device=dev)
                xp = ops.patch_gather(slab, W0, idx=both, noise=z, noise_scale=0.5)      # train.py:157,170 fused
                xs = spec[both] + torch.randn((256, B0), device=dev) * 0.5              # train.py:158,171
                return xp, xs
            xb, sb = batch()
            xe, se = batch()
            T.mutual_step(tst, xb, sb, xe, se, truth[li], 1, i, targs)
        for i in range(3):
            train_step(i)
        t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t0e.record()
        for i in range(10):
            train_step(3 + i)
        t1e.record()
        torch.cuda.synchronize()
        train_ms = t0e.elapsed_time(t1e) / 10

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = measured_peaks()
    px_step = R0 * C0 * world
    value = px_step / (ms_dev / 1e3)
    e2e_val = px_step / (ms_e2e / 1e3)
    # dominant kernel: conv2_scene (tensor-bound).  FLOPs it EXECUTES: 169 tap products (64x64 MACs) per plane
    # position over the 25 border classes; the reference's per-patch arithmetic for the same layer is
    # FLOP_PER_PX_CONV2 per pixel -- reported next to it, it exceeds the hardware peak because of the sharing.
    cnn_ms = stage_ms["conv2_scene"]
    qpos = 4 * ((nb + W0) // 2) * ((C0 + W0) // 2)
    conv2_exec_flop = qpos * 169 * 2 * 64 * 64
    achieved = conv2_exec_flop / (cnn_ms / 1e3) / 1e12
    ppos_n = (nb + W0 - 1) * (C0 + W0 - 1)
    conv1_exec_flop = ppos_n * 2 * 64 * 64 * 9           # 9 single-tap products per position (column classes summed in the epilogue)
    traffic = None
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get("conv2_scene_dram_bytes_per_launch")
    # conv0_map, x16_tile, spectral_logits, conv1_pool, conv2_scene, pool2_cls, head_sum (+ confusion when sharded)
    launches_per_step = 7 + (1 if world > 1 else 0)
    line = {
        "metric": "pixels/sec full-scene inference", "value": value, "unit": "pixels/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None,
        "dtype": "f16 operands / f32 accumulate (tcgen05 kind::f16; same 11-bit significand as the TF32 the "
                 "reference's torch 1.8 GPU path used); fp32 elsewhere",
        "data": "synthetic", "config": workload_config(world),
        "clocks": clocks,
        "e2e": {"value": e2e_val, "unit": "pixels/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": int(raw_host.numel() * 2) * world,
                "d2h_bytes_per_step": int(n_band) * world,
                "input": "raw uint16 cube (pinned host), preprocessing parameters fitted beforehand; z-score + PCA "
                         "projection applied on device inside the timed region, folded into conv0 and into the fp16 "
                         "conversion of the spectra (cmlpl_scene_infer_raw)",
                "f32_inputs": {"value": px_step / (ms_e2e_f32 / 1e3), "ms_per_step": ms_e2e_f32,
                               "h2d_bytes_per_step": int(slab_host.numel() * 4 + spec_host.numel() * 4) * world,
                               "input": "already preprocessed f32 PCA cube + f32 spectra (pinned host)"}},
        "gpu_launches": launches_per_step * args.steps,
        "roofline": {"kernel": "conv2_scene_kernel (tcgen05 conv2 + residual + ReLU once per scene position in 25 patch-border "
                               "classes, parity planes, row-tap fusion into N=192/128 MMAs, TMA tile loads)", "bound": "tensor",
                     "achieved": achieved, "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                     "frac": achieved / peaks["tf_sustained"], "traffic": traffic,
                     "peak_source": f"{peaks['src']} bf16 sustained (kernel timed inside the step loop)",
                     "executed_flop_per_launch": conv2_exec_flop, "positions_per_launch": qpos,
                     "pixels_per_launch": n_band,
                     "reference_arithmetic_tflops": n_band * FLOP_PER_PX_CONV2 / (cnn_ms / 1e3) / 1e12,
                     "note": "exact compute sharing (SURVEY section 7 / 8-f3): conv1 and conv2 are evaluated once per scene "
                             "position in 9 / 25 patch-border classes instead of once per pixel patch, so `achieved` counts the "
                             "useful FLOPs the kernel executes (169 tap products of 64x64 MACs per position); in the reference's per-patch arithmetic (7.37 MFLOP/pixel for conv2, "
                             "40.2 for the net) the same launch / step is reference_arithmetic_tflops / "
                             "whole_step_algorithmic_tflops, above the hardware peak by the sharing factor",
                     "conv1_pool_executed_tflops": conv1_exec_flop / (stage_ms["conv1_pool"] / 1e3) / 1e12,
                     "kernel_ms": cnn_ms, "stage_ms": stage_ms,
                     "whole_step_algorithmic_tflops": px_step / world * FLOP_PER_PX_ALL / (ms_dev / 1e3) / 1e12},
        "host_prep_s": t_data,
        "train_step_ms": train_ms,
        "train_step_config": "BaseNet2 x2 mutual-learning step (train.py:149-278), 128 labelled + 128 unlabelled, "
                             "dropout 0.8, noise 0.5 from the device generator, smoothing branch on; fp32 kernels",
    }
    if gather is not None:
        gather["achieved"] = 16384 * 96240 / (gather["ms"] / 1e3) / 1e9
        gather["peak"] = peaks["hbm"]
        gather["unit"] = "GB/s"
        gather["frac"] = gather["achieved"] / peaks["hbm"]
        gather["frac_random_pixels"] = 16384 * 96240 / (gather["ms_random_pixels"] / 1e3) / 1e9 / peaks["hbm"]
        line["aux_roofline"] = gather
    if world == 1 and not args.no_cpu_baseline:
        from oracle import cmlpl_oracle as O      # cpu_baseline leg: the oracle port is the thing timed here
        torch.manual_seed(1088)
        sd = O.basenet2_init(B0, K0)
        threads = os.cpu_count() or 1
        tot_t, tot_n, rows = 0.0, 0, 6
        rr = 100
        while tot_t < 12.0 and rr + rows < R0:
            dt, n = cpu_reference_pass(cube_pca, spectra, sd, (rr, rr + rows), threads)
            tot_t += dt; tot_n += n; rr += rows
        line["cpu_baseline"] = {"value": tot_n / tot_t, "unit": "pixels/s", "cores": threads, "kind": "port",
                                "sample": f"{tot_n} px (rows 100..{rr} of the same scene) in {tot_t:.1f} s: oracle port "
                                          "of the reference test_whole path (per-pixel patch loop + BaseNet2 fp32 bs 512 "
                                          "+ argmax), torch threads = all host cores"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
