"""Scene preprocessing on the GPU: the reference's ``featureNormalize`` + ``PCANorm``
(tools/hyper_tools.py:8-32, used at :289-292) as fit (float64 moments on device, B x B SVD on the
host) and apply (raw cube -> z-scored spectra + z-scored PCA cube in one pass).  Both inputs of the
scene path are affine functions of the raw cube, so ``apply`` lets the end-to-end entry point ship only
the raw uint16 cube over PCIe."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from . import _lib


@dataclass
class Preproc:
    mu: np.ndarray          # f64 [B]   band means                         (featureNormalize / PCANorm mean)
    sigma: np.ndarray       # f64 [B]   band population std                (np.std, hyper_tools.py:14)
    U: np.ndarray           # f64 [B, n_PC] leading left-singular vectors of np.cov   (hyper_tools.py:29-31)
    pca_mu: np.ndarray      # f64 [n_PC] mean of the projected data (~0)
    pca_sigma: np.ndarray   # f64 [n_PC] population std of the projected data
    _dev: dict = None

    def device_params(self, device):
        if self._dev is None or self._dev["device"] != device:
            t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(device)
            self._dev = {"device": device, "mu": t(self.mu), "inv_sigma": t(1.0 / self.sigma),
                         "Us": t(self.U / self.pca_sigma[None, :]), "shift": t(self.pca_mu / self.pca_sigma)}
        return self._dev

    def folded_conv0(self, conv0_w, conv0_b, device):
        """PCA projection, both z-scores and conv0 (models.py:102,132) are per-pixel affine maps of the raw
        spectrum, so they fold into F0 = wf^T (x - mu) + bf.  Returns device f32 tensors
        wf [B, 64], bf [64] (+ mu, inv_sigma [B]) for cmlpl_scene_infer_raw; folded in float64 on the host."""
        W0 = conv0_w.detach().double().cpu().numpy().reshape(conv0_w.shape[0], -1)        # [64, n_PC]
        b0 = conv0_b.detach().double().cpu().numpy()
        Us = self.U / self.pca_sigma[None, :]                                              # [B, n_PC]
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(device)
        d = self.device_params(device)
        return {"wf": t(Us @ W0.T), "bf": t(b0 - W0 @ (self.pca_mu / self.pca_sigma)), "mu": d["mu"],
                "inv_sigma": d["inv_sigma"]}


def _dtype_code(x: torch.Tensor) -> int:
    if x.dtype == torch.uint16:
        return 0
    if x.dtype == torch.float32:
        return 1
    raise _lib.CmlplError(f"raw scene must be uint16 or float32, got {x.dtype}")


def fit(raw: torch.Tensor, n_PC: int = 60) -> Preproc:
    """raw: CUDA tensor [N, B] (uint16 or float32).  Moments in float64 on device, SVD on the host."""
    if not raw.is_cuda or not raw.is_contiguous():
        raise _lib.CmlplError("fit needs a contiguous CUDA tensor")
    n, B = raw.shape
    mean = torch.empty(B, dtype=torch.float64, device=raw.device)
    gram = torch.empty(B, B, dtype=torch.float64, device=raw.device)
    _lib.call("cmlpl_preprocess_fit_f64", raw.data_ptr(), _dtype_code(raw), n, B, mean.data_ptr(), gram.data_ptr(),
              torch.cuda.current_stream().cuda_stream)
    g = gram.cpu().numpy()
    g = np.triu(g) + np.triu(g, 1).T                              # the kernel fills upper tiles only
    mu = mean.cpu().numpy()
    U = np.linalg.svd(g / (n - 1))[0][:, :n_PC]                   # np.cov + np.linalg.svd (hyper_tools.py:29-30)
    pca_var = np.einsum("bk,bc,ck->k", U, g, U) / n               # population variance of the projection
    return Preproc(mu=mu, sigma=np.sqrt(np.diag(g) / n), U=U, pca_mu=np.zeros(n_PC), pca_sigma=np.sqrt(pca_var))


def apply(raw: torch.Tensor, pp: Preproc, want_spectra: bool = True, cube: torch.Tensor | None = None,
          spectra: torch.Tensor | None = None):
    """raw CUDA [N, B] -> (cube f32 [N, n_PC], spectra f32 [N, B] or None)."""
    if not raw.is_cuda or not raw.is_contiguous():
        raise _lib.CmlplError("apply needs a contiguous CUDA tensor")
    n, B = raw.shape
    d = pp.device_params(raw.device)
    npc = d["Us"].shape[1]
    if cube is None:
        cube = torch.empty((n, npc), dtype=torch.float32, device=raw.device)
    if want_spectra and spectra is None:
        spectra = torch.empty((n, B), dtype=torch.float32, device=raw.device)
    _lib.call("cmlpl_preprocess_apply", raw.data_ptr(), _dtype_code(raw), n, B, npc, d["mu"].data_ptr(),
              d["inv_sigma"].data_ptr(), d["Us"].data_ptr(), d["shift"].data_ptr(), cube.data_ptr(),
              spectra.data_ptr() if want_spectra else None, torch.cuda.current_stream().cuda_stream)
    return cube, (spectra if want_spectra else None)
