"""The intermediate tensors of the dense scene path, stage by stage (tools/models.py:137-150 for all pixels at once):
the half-pooled conv2 maps  YP[Al][Be]  that conv2_scene_kernel writes (border halves of the 2x2 pool added in its
epilogue) and the row maps  M[I] = sum_J L[I][J](., x + 2J)  that pool2_cls_kernel writes, each against a plain torch
fp32 evaluation of the same formula on the kernel's own fp16 inputs, then the whole path against the independent
per-pixel kernels and the CPU oracle.  The checker is scripts/check_dense_path.py (also run by hand with --big for
the PaviaU-size stage timings).  Bars: fp16-rounded maps 1e-3 * max|ref| (one rounding, 4.9e-4 relative), fp32 class
partial maps 1e-4, logits 1e-3 * max|ref| (north_star), labels identical on these sizes."""
import importlib.util
import os

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _checker():
    spec = importlib.util.spec_from_file_location("check_dense_path", os.path.join(ROOT, "scripts", "check_dense_path.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("R,C,B,K,seed", [(37, 45, 103, 9, 1), (24, 75, 144, 15, 2), (50, 61, 200, 16, 5)])
def test_dense_stages_match_torch_and_oracle(dev, R, C, B, K, seed):
    res = _checker().run(R, C, B, K, seed)
    assert res["planes_equal"]
    assert res["half_pooled_nan"] == 0 and res["half_pooled_rel"] < 1e-3, res
    assert res["row_maps_rel"] < 1e-4, res
    assert res["dense_vs_pixel_rel"] < 1e-3 and res["labels_equal"] == 1.0, res
    assert res["dense_vs_oracle_rel"] < 1e-3, res
