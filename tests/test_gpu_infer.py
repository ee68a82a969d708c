"""Sliding-window inference (segtran_b200.inference.test_single_case, csrc/sx_infer.cu) against the reference-generated
fixtures (masks exact wherever the soft prediction is not within rounding of the 0.5 threshold) and, for the padded
case the reference cannot run, against the CPU oracle."""
import pytest
import torch

from oracle import infer_oracle as IO
from tests.helpers import AffinePickNet, load_golden

pytestmark = pytest.mark.gpu


def _check(hard, soft, ref_hard, ref_soft, brats):
    soft, hard = soft.cpu(), hard.cpu()
    assert soft.shape == ref_soft.shape and hard.shape == ref_hard.shape
    assert (soft - ref_soft).abs().max() < 2e-6
    if brats:
        sure = ((ref_soft - 0.5).abs() > 1e-5)                     # away from the threshold the hard masks must be identical
        sure[0] = sure[1:].all(dim=0)
        assert torch.equal(hard[sure], ref_hard[sure])
        assert float(sure.float().mean()) > 0.999
    else:
        top2 = ref_soft.topk(2, dim=0).values
        sure = (top2[0] - top2[1]) > 1e-5
        assert torch.equal(hard[sure], ref_hard[sure])


def test_sliding_window_matches_reference_fixtures():
    from segtran_b200.inference import test_single_case
    fx = load_golden("infer_sw")
    for key, c in fx["cases"].items():
        net = AffinePickNet(c["a"], c["b"], c["ch"]).cuda()
        hard, soft = test_single_case(net, c["image"].cuda(), c["orig_patch"], c["input_patch"], c["batch_size"],
                                      c["stride_xy"], c["stride_z"], c["task"], "segtran", c["K"])
        _check(hard, soft, c["hard"], c["soft"], c["task"] == "brats")
        assert hard.dtype == c["hard"].dtype


def test_sliding_window_padded_volume_matches_oracle():
    from segtran_b200.inference import test_single_case
    torch.manual_seed(3)
    image = torch.randn(4, 20, 30, 12) * 2
    net = AffinePickNet([1.0, 1.5, 0.7, 1.2], [0.1, -0.2, 0.3, 0.0], [0, 1, 2, 3])
    args = ((24, 24, 16), (24, 24, 16), 2, 12, 8, "brats", "segtran", 4)
    ref_hard, ref_soft = IO.test_single_case(net, image, *args)
    hard, soft = test_single_case(net.cuda(), image.cuda(), *args)
    assert hard.shape == (4, 20, 30, 12)
    _check(hard, soft, ref_hard, ref_soft, True)
