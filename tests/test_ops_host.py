"""Host-side logic of segtran_b200.ops that needs no GPU: the per-call-site precision policy and the
gradient-ready milestone hooks (autograd runs nodes in descending creation order, so a hook on a tensor fires after every
node created later — the property the milestone all-reduce relies on)."""
import pytest
import torch

from segtran_b200 import ops


def test_small_tag_follows_the_bank_to_token_ratio_of_the_baseline_configs():
    # (attractors, tokens) of BASELINE configs 1-5: the 2-D configs run their attractor-row products 3-pass, cfg 4/5 do not
    assert ops.small_tag(256, 1296) == "small"
    assert ops.small_tag(256, 5184) == "small"
    assert ops.small_tag(256, 1936) == "small"
    assert ops.small_tag(1024, 2744) == "smallwide"
    assert ops.small_tag(2048, 5832) == "smallwide"
    assert ops.small_tag(256, 1024) == "small" and ops.small_tag(257, 1024) == "smallwide"      # boundary: 4 A <= N


def test_precision_policy_switches_and_rounding_flags():
    saved = ops.get_precision_policy()
    try:
        assert saved == {"small": "tf32x3", "smallwide": "tf32", "proj": "tf32", "insq": "tf32", "big": "tf32"}
        # a producer feeding ONLY 3-pass contractions must not round its output to TF32 (the split needs the full mantissa)
        assert ops.rt_for("small") == 0 and ops.rt_for("big") == 1 and ops.rt_for("smallwide") == 1
        ops.set_precision_policy(small="tf32", proj="tf32x3")
        assert ops.rt_for("small") == 1 and ops.rt_for("proj") == 0
        with pytest.raises(ValueError):
            ops.set_precision_policy(big="bf16")
        with pytest.raises(ValueError):
            ops.set_precision_policy(huge="tf32")
        ops.set_precision("tf32x3")                  # validation mode: nothing is rounded, whatever the policy says
        assert all(ops.rt_for(t) == 0 for t in saved)
        ops.set_precision("bf16")                    # bf16 mode: the policy does not apply (no 3-pass products)
        assert all(ops.rt_for(t) == 1 for t in saved)
        with pytest.raises(ValueError):
            ops.set_precision("fp8")
    finally:
        ops.set_precision("tf32")
        ops.set_precision_policy(**saved)


def test_grad_ready_hooks_fire_in_reverse_layer_order_with_final_gradients():
    """Two 'layers'; a milestone on the tensor between them must fire after layer 2's parameter gradient is final and
    before layer 1's backward has produced its own."""
    w1 = torch.nn.Parameter(torch.randn(4, 4))
    w2 = torch.nn.Parameter(torch.randn(4, 4))
    events = []

    def cb(params):
        events.append([None if p.grad is None else p.grad.clone() for p in (w1, w2)])

    ops.set_grad_ready_callback(cb)
    try:
        x0 = torch.randn(3, 4, requires_grad=True)
        x = x0 * 1.0                                 # the marked tensors are results of operations, as in the modules
        h = x @ w1
        ops.grad_ready(h, [w2])                      # everything after h uses w2 only
        y = (h @ w2).sum()
        ops.grad_ready(x, [w1])
        ops.grad_ready(x0, [w1])                     # a leaf is not marked (the order of sibling AccumulateGrads is unspecified)
        y.backward()
    finally:
        ops.set_grad_ready_callback(None)
    assert len(events) == 2
    g1_at_h, g2_at_h = events[0]
    assert g1_at_h is None                           # layer 1 has not run its backward yet ...
    assert torch.equal(g2_at_h, w2.grad)             # ... while layer 2's gradient is already final
    assert torch.equal(events[1][0], w1.grad)        # at the layer input, layer 1's gradient is final too
    # without a callback the marker is a no-op
    ops.grad_ready(torch.randn(2, requires_grad=True) * 2, [w1])
