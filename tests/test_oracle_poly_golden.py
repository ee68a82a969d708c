"""The Polyformer restatement in oracle/segtran_oracle.py against the reference-generated fixture."""
import torch

from oracle import segtran_oracle as O
from tests.helpers import load_golden, rel_err


def test_polyformer_oracle_matches_reference_fixture():
    fx = load_golden("poly2d_tiny")
    p = {k: v for k, v in fx["state_dict"].items()}
    with torch.no_grad():
        y = O.polyformer_layer(p, "polyformer_layers.0.", fx["x"], fx["args"]["num_modes"])
    assert rel_err(y, fx["out"]) < 1e-6
