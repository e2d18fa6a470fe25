"""Edge cases of the scene path on GPU (-m gpu): tiny / ragged scenes, odd pixel counts, single-row
bands, class/band counts that take the CUDA-core head, a Houston-shaped scene checked on sampled rows."""
import numpy as np
import pytest
import torch

from oracle import cmlpl_oracle as O

pytestmark = pytest.mark.gpu


def rel(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def _run(dev, R, C, B, K, seed, **kw):
    from cmlpl_b200 import ops
    rng = np.random.default_rng(seed)
    cube = rng.standard_normal((R, C, 60)).astype(np.float32)
    spectra = rng.standard_normal((R * C, B)).astype(np.float32)
    torch.manual_seed(seed)
    sd = O.basenet2_init(B, K)
    packed = ops.pack_basenet2({k: v.to(dev) for k, v in sd.items()}, B, K, 20)
    labels, logits = ops.scene_infer(torch.from_numpy(cube).to(dev), torch.from_numpy(spectra).to(dev), packed, K, 20,
                                     want_logits=True, **kw)
    return sd, cube, spectra, labels.cpu().numpy(), logits.cpu().numpy()


@pytest.mark.parametrize("R,C", [(10, 10), (11, 13), (10, 31), (33, 10)])
def test_tiny_and_odd_scenes(dev, R, C):
    """Smallest scenes the mirror padding allows (w/2 <= min dim), odd pixel counts (pair tail)."""
    sd, cube, spectra, lab, logits = _run(dev, R, C, 103, 9, seed=R * 100 + C)
    lab_ref, log_ref = O.test_whole(sd, cube, spectra, 20, return_logits=True)
    assert rel(logits, log_ref) < 1e-3
    assert np.array_equal(lab, logits.argmax(1))
    assert np.mean(lab == lab_ref) >= 0.97


def test_single_row_bands_are_bit_identical(dev):
    from cmlpl_b200 import ops
    R, C, B, K = 14, 21, 103, 9
    rng = np.random.default_rng(2)
    cube = torch.from_numpy(rng.standard_normal((R, C, 60)).astype(np.float32)).to(dev)
    spectra = torch.from_numpy(rng.standard_normal((R * C, B)).astype(np.float32)).to(dev)
    torch.manual_seed(2)
    packed = ops.pack_basenet2({k: v.to(dev) for k, v in O.basenet2_init(B, K).items()}, B, K, 20)
    full = ops.scene_infer(cube, spectra, packed, K, 20)
    rows = [ops.scene_infer(cube, spectra[r * C:(r + 1) * C].contiguous(), packed, K, 20, band_row0=r, band_rows=1)
            for r in range(R)]
    assert torch.equal(torch.cat(rows), full)


@pytest.mark.parametrize("B,K", [(224, 16), (103, 20), (40, 3)])
def test_cuda_core_head_and_unusual_widths(dev, B, K):
    """> 208 bands or > 16 classes take the all-per-pixel kernel + CUDA-core head; small B/K the tensor path."""
    sd, cube, spectra, lab, logits = _run(dev, 24, 26, B, K, seed=B + K)
    lab_ref, log_ref = O.test_whole(sd, cube, spectra, 20, return_logits=True)
    assert rel(logits, log_ref) < 1e-3
    assert np.array_equal(lab, logits.argmax(1))


def test_houston_shaped_scene_sampled_rows(dev):
    """BASELINE.json configs[3] shape (349 x 1905 x 144, 15 classes): whole scene on GPU, oracle on sampled rows."""
    R, C, B, K = 349, 1905, 144, 15
    sd, cube, spectra, lab, logits = _run(dev, R, C, B, K, seed=4)
    assert lab.shape == (R * C,) and np.array_equal(lab, logits.argmax(1))
    for r in (0, 173, 348):                                  # first, middle, last row (mirror at both edges)
        cols = np.arange(0, C, 37)
        idx = r * C + cols
        XP = O.extract_patches_at(cube, 20, idx)
        with torch.no_grad():
            ref, _ = O.basenet2_forward(sd, torch.from_numpy(XP), torch.from_numpy(spectra[idx]))
        assert rel(logits[idx], ref.numpy()) < 1e-3, r


def test_confusion_and_gather_empty_inputs(dev):
    from cmlpl_b200 import ops
    cm = ops.confusion(torch.zeros(0, dtype=torch.uint8, device=dev), torch.zeros(0, dtype=torch.int64, device=dev), 5)
    assert int(cm.sum()) == 0
    lab = torch.tensor([0, 1, 2, 250], dtype=torch.uint8, device=dev)
    tru = torch.tensor([0, -1, 2, 3], dtype=torch.int64, device=dev)           # -1 (unlabelled) and out-of-range ignored
    cm = ops.confusion(lab, tru, 4).cpu().numpy()
    assert cm.sum() == 2 and cm[0, 0] == 1 and cm[2, 2] == 1
