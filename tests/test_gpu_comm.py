"""The collective entry points of the C ABI (cmlpl_comm_*, NCCL resolved inside libcmlpl_sm100.so): a one-rank
communicator on any GPU box, and two ranks (one process per GPU) when the box has two GPUs -- the band-sharded label
map and confusion matrix equal the single-GPU ones."""
import os

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def test_single_rank_communicator(dev):
    from cmlpl_b200.parallel import CComm
    c = CComm(0, 1, CComm.unique_id())
    lab = torch.arange(0, 200, dtype=torch.int64, device=dev).to(torch.uint8)
    out = c.gather_label_map(lab, 10, 20)
    cm = torch.arange(81, dtype=torch.int64, device=dev).view(9, 9).contiguous()
    want = cm.clone()
    c.reduce_confusion(cm)
    torch.cuda.synchronize()
    assert torch.equal(out, lab) and torch.equal(cm, want)
    c.close()


def _rank(rank, world, id_path, out_path):
    import time
    torch.cuda.set_device(rank)
    from cmlpl_b200 import ops
    from cmlpl_b200.parallel import CComm, band_of, slab_of
    from cmlpl_b200.tools.models import BaseNet2
    if rank == 0:
        with open(id_path + ".tmp", "wb") as f:
            f.write(CComm.unique_id())
        os.replace(id_path + ".tmp", id_path)
    while not os.path.exists(id_path):
        time.sleep(0.05)
    uid = open(id_path, "rb").read()
    c = CComm(rank, world, uid)
    R, C, B, K = 37, 29, 103, 9
    rng = np.random.default_rng(5)
    cube = rng.standard_normal((R, C, 60)).astype(np.float32)
    spectra = rng.standard_normal((R * C, B)).astype(np.float32)
    truth = torch.from_numpy(rng.integers(0, K, R * C)).cuda()
    torch.manual_seed(1088)
    net = BaseNet2(B, 0, K).cuda().eval()
    packed = net.packed_weights(20)
    r0, r1 = band_of(rank, world, R)
    s0, s1 = slab_of(r0, r1, R, 20)
    lab = ops.scene_infer(torch.from_numpy(cube[s0:s1]).cuda(), torch.from_numpy(spectra[r0 * C:r1 * C]).cuda(), packed, K, 20,
                          band_row0=r0, band_rows=r1 - r0, scene_rows=R, slab_row0=s0)
    full = c.gather_label_map(lab, R, C)
    cm = ops.confusion(lab, truth[r0 * C:r1 * C].contiguous(), K)
    c.reduce_confusion(cm)
    torch.cuda.synchronize()
    if rank == 0:
        one = ops.scene_infer(torch.from_numpy(cube).cuda(), torch.from_numpy(spectra).cuda(), packed, K, 20)
        cm1 = ops.confusion(one, truth, K)
        ok = bool(torch.equal(full, one) and torch.equal(cm, cm1))
        open(out_path, "w").write("ok" if ok else "mismatch")
    c.close()


def test_two_rank_band_sharding_through_the_c_abi(dev, tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    id_path, out_path = str(tmp_path / "nccl_id"), str(tmp_path / "result")
    mp.spawn(_rank, args=(2, id_path, out_path), nprocs=2, join=True)
    assert open(out_path).read() == "ok"
