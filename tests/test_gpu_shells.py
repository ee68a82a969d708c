"""GPU parity of the Segtran3d / Segtran2d shells (flatten -> stack -> scatter -> collapsed head) against golden
fixtures produced by the real reference shells with a fixed-feature backbone (oracle/gen_golden.py)."""
from argparse import Namespace

import pytest
import torch

from tests.helpers import load_golden, rel_err

pytestmark = pytest.mark.gpu

OUT_TOL = 1e-3      # logits, max|a-b|/max|b|  (north_star)
GRAD_TOL = 5e-3


class FixedFeat3d(torch.nn.Module):
    def __init__(self, feats):
        super().__init__()
        self.feats = feats

    def extract_features(self, x):
        keys = ["MaxPool3d_2a_3x3", "Conv3d_2c_3x3", "Mixed_3c", "Mixed_4f", "Mixed_5c"]
        return dict(zip(keys, self.feats))


class FixedFeat2d(torch.nn.Module):
    def __init__(self, feats):
        super().__init__()
        self.feats = feats

    def ext_features(self, x):
        return tuple(self.feats)


def _build(kind):
    import segtran_b200.networks.segtran_shared as S
    fx = load_golden("seg3d_tiny" if kind == 3 else "seg2d_tiny")
    args = Namespace(**fx["args"])
    args.device = "cuda"
    S.bb2feat_dims[args.backbone_type] = fx["bb_feat_dims"]
    feats = [f.cuda().requires_grad_(i > 0) for i, f in enumerate(fx["feats"])]
    if kind == 3:
        import segtran_b200.networks.segtran3d as M
        cfg = M.Segtran3dConfig()
        cfg.update_config(args)
        net = M.Segtran3d(cfg, backbone=FixedFeat3d(feats))
    else:
        import segtran_b200.networks.segtran2d as M
        cfg = M.Segtran2dConfig()
        cfg.update_config(args)
        net = M.Segtran2d(cfg, backbone=FixedFeat2d(feats))
    missing, unexpected = net.load_state_dict(fx["state_dict"], strict=False)
    assert not unexpected and not missing, (missing, unexpected)
    return fx, net.cuda().eval(), feats


@pytest.mark.parametrize("kind", [3, 2])
def test_shell_logits_and_masks(kind):
    fx, net, feats = _build(kind)
    with torch.no_grad():
        y = net(fx["batch"].cuda())
    assert y.shape == fx["out"].shape
    e = rel_err(y, fx["out"])
    print("seg%dd logits rel" % kind, e)
    assert e < OUT_TOL
    # hard masks (sigmoid >= 0.5 <=> logit >= 0) must agree except where the reference logit is within the tolerance of 0
    ref = fx["out"]
    safe = ref.abs() > OUT_TOL * ref.abs().max()
    assert torch.equal((y.cpu() >= 0)[safe], (ref >= 0)[safe])


@pytest.mark.parametrize("kind", [3, 2])
def test_shell_gradients(kind):
    fx, net, feats = _build(kind)
    y = net(fx["batch"].cuda())
    (y * fx["G"].cuda()).sum().backward()
    for i in range(1, 5):
        e = rel_err(feats[i].grad, fx["grad_feats"][i])
        print("seg%dd dfeat%d rel" % (kind, i), e)
        assert e < GRAD_TOL, i
    gscale = max(float(g.abs().max()) for g in fx["grad_params"].values())
    got = dict(net.named_parameters())
    for k, g in fx["grad_params"].items():
        if k.startswith("backbone."):
            continue
        gg = got[k].grad
        if float(g.abs().max()) == 0.0:
            assert gg is None or float(gg.abs().max()) <= 1e-5 * gscale, k
            continue
        assert gg is not None, k
        err = float((gg.cpu() - g).abs().max())
        assert err <= GRAD_TOL * float(g.abs().max()) + 2e-5 * gscale, (k, err, float(g.abs().max()))
