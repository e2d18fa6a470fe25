"""Host logic of the row-band sharding (cmlpl_b200/parallel.py) on CPU: gloo, world_size 2 and 3.
The per-band compute is stood in by the CPU oracle (the CUDA path is covered by -m gpu tests);
what is checked here is the partition, the halo slabs and the two collectives."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cmlpl_b200 import parallel
from oracle import cmlpl_oracle as O


def test_band_and_slab_match_oracle():
    for R in (13, 37, 349, 610):
        for world in (1, 2, 3, 4, 8):
            covered = []
            for rank in range(world):
                r0, r1 = parallel.band_of(rank, world, R)
                s0, s1 = parallel.slab_of(r0, r1, R, 20)
                o = O.band_rows(R, world, rank, 20)
                assert (r0, r1) == (o[0], o[1])
                if r1 > r0:
                    assert (s0, s1) == (o[2], o[3]), (R, world, rank)
                covered += list(range(r0, r1))
            assert covered == list(range(R))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, R, C, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(0)
        cube = rng.standard_normal((R, C, 60)).astype(np.float32)
        spectra = rng.standard_normal((R * C, 16)).astype(np.float32)
        truth = torch.from_numpy(rng.integers(0, 5, size=R * C))
        torch.manual_seed(0)
        sd = O.basenet2_init(16, 5)

        def infer_band(r0, r1):
            if r1 <= r0:
                return torch.zeros(0, dtype=torch.uint8)
            s0, s1 = parallel.slab_of(r0, r1, R, 20)
            idx = np.arange(r0 * C, r1 * C)
            XP = O.extract_patches_at(cube[s0:s1], 20, idx, scene_rows=R, row0=s0)   # halo-only slab
            with torch.no_grad():
                lo, _ = O.basenet2_forward(sd, torch.from_numpy(XP), torch.from_numpy(spectra[idx]))
            return lo.argmax(1).to(torch.uint8)

        conf = lambda pred, lab, K: torch.from_numpy(O.confusion_matrix(pred.numpy(), lab.numpy(), K))
        labels, cm = parallel.sharded_scene_labels(infer_band, R, C, truth, 5, conf)
        if rank == 0:
            ref = O.test_whole(sd, cube, spectra, 20)
            ok = np.array_equal(labels.numpy().astype(np.int64), ref)
            ok &= np.array_equal(cm.numpy(), O.confusion_matrix(ref, truth.numpy(), 5))
            out.put(bool(ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,R", [(2, 23), (3, 22)])
def test_sharded_inference_gloo(world, R):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, R, 21, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    assert out.get(timeout=5) is True


def _worker_exchange(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        R, C, K = 11, 7, 5
        rng = np.random.default_rng(3)
        full = torch.from_numpy(rng.integers(0, K, R * C).astype(np.uint8))
        truth = torch.from_numpy(rng.integers(0, K, R * C))
        r0, r1 = parallel.band_of(rank, world, R)
        local = full[r0 * C:r1 * C]
        cm_local = torch.from_numpy(O.confusion_matrix(local.numpy(), truth[r0 * C:r1 * C].numpy(), K))
        ex = parallel.SceneExchange(R, C, K, torch.device("cpu"))
        for _ in range(2):                                   # buffers are reused across calls
            labels, cm = ex(local, cm_local)
        if rank == 0:
            ok = torch.equal(labels, full) and np.array_equal(cm.numpy(), O.confusion_matrix(full.numpy(), truth.numpy(), K))
            out.put(bool(ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_scene_exchange_single_collective(world):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_exchange, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out.get(timeout=5) is True
