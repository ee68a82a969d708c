"""The inference-path oracle (oracle/infer_oracle.py) against fixtures made by the reference's own
test_util3d.test_single_case (tests/golden/infer_sw.pt, oracle/gen_golden.py infer)."""
import torch

from oracle import infer_oracle as IO
from tests.helpers import AffinePickNet, load_golden


def test_infer_oracle_matches_reference_fixtures():
    fx = load_golden("infer_sw")
    for key, c in fx["cases"].items():
        net = AffinePickNet(c["a"], c["b"], c["ch"])
        hard, soft = IO.test_single_case(net, c["image"], c["orig_patch"], c["input_patch"], c["batch_size"], c["stride_xy"],
                                         c["stride_z"], c["task"], "segtran", c["K"])
        assert torch.equal(soft, c["soft"]), key
        assert torch.equal(hard, c["hard"]), key
