"""Parity at BASELINE.json's full cfg-4 sizes (N=2744 tokens, C=F=1024, A=1024 attractors, M=4), one sample:
the CUDA stack + collapsed head against the CPU oracle (reference formulation, un-collapsed head on a spatial crop),
plus size-independent properties at the full batch: per-sample independence and linearity of the head."""
import pytest
import torch

from oracle import segtran_oracle as O
from tests.helpers import encoder_config, rel_err, rms_rel

pytestmark = pytest.mark.gpu


def _encoder(dims, A, M=4, seed=0):
    import segtran_b200.networks.segtran_shared as S
    cfg = encoder_config(S.SegtranConfig, dims=dims, num_modes=M, num_attractors=A, pos_dim=3)
    torch.manual_seed(seed)
    enc = S.SegtranFusionEncoder(cfg, "Fusion")
    init = S.SegtranInitWeights(cfg)
    enc.apply(init.init_weights)
    enc.apply(init.tie_qk)
    enc.apply(init.add_identity_bias)
    return enc


def test_cfg4_stack_matches_oracle_one_sample():
    dims, A, grid = [1024, 1024], 1024, (14, 14, 14)
    enc = _encoder(dims, A).eval()
    p = {"voxel_fusion." + k: v.clone() for k, v in enc.state_dict().items()}
    torch.manual_seed(1)
    x = torch.randn(1, 2744, 1024) * 1.5 + 0.2
    pos = O.voxels_pos_for_grid(grid, (8, 8, 8), 1)
    mask = torch.ones(1, 2744, 1)
    with torch.no_grad():
        ref = O.fusion_encoder(p, "voxel_fusion.", x, pos, mask, dims, 4)
        y = enc.cuda()(x.cuda(), pos.cuda(), mask.cuda(), torch.Size(grid))
    e, r = rel_err(y, ref), rms_rel(y, ref)
    print("cfg4 stack: max-rel %.3e rms-rel %.3e" % (e, r))
    assert e < 1e-3          # north_star tolerance, max|a-b|/max|b|


def test_cfg4_batch_independence_and_determinism():
    """Every op on the path is per-sample: sample 0 of a batch of 4 equals the same sample run alone (bit-exact,
    eval mode), and two runs of the same batch are identical."""
    enc = _encoder([1024, 1024], 1024).cuda().eval()
    torch.manual_seed(2)
    x = torch.randn(4, 2744, 1024, device="cuda")
    pos = O.voxels_pos_for_grid((14, 14, 14), (8, 8, 8), 1).cuda().expand(4, -1, -1)
    with torch.no_grad():
        y4 = enc(x, pos, None, torch.Size((14, 14, 14)))
        y4b = enc(x, pos, None, torch.Size((14, 14, 14)))
        y1 = enc(x[:1].contiguous(), pos[:1], None, torch.Size((14, 14, 14)))
    assert torch.equal(y4, y4b)
    assert rel_err(y4[:1], y1) < 1e-6


def test_cfg4_head_linearity_and_crop_parity():
    """Head at the full 56^3 / 112^3 sizes: linear in (curr, vfeat) jointly (a size-independent property), and equal
    to the oracle's un-collapsed head on a problem small enough for the CPU (same channel widths)."""
    from segtran_b200 import ops
    torch.manual_seed(3)
    B, Cf, Fd, K = 1, 832, 1024, 4
    Wb = torch.randn(Fd, Cf, 1, 1, 1) * 0.03
    bb = torch.randn(Fd) * 0.1
    Wc = torch.randn(K, Fd, 1, 1, 1) * 0.03
    bc = torch.randn(K) * 0.1
    W = [t.cuda() for t in (Wb, bb, Wc, bc)]

    def head(curr, vf, grid, out):
        return ops.seg_head(curr, vf, grid, W[0], W[1], W[2], W[3], out, d_pool_k=2)

    c1 = torch.randn(B, Cf, 56, 56, 56, device="cuda")
    c2 = torch.randn(B, Cf, 56, 56, 56, device="cuda")
    v1 = torch.randn(B, 2744, Fd, device="cuda")
    v2 = torch.randn(B, 2744, Fd, device="cuda")
    with torch.no_grad():
        y1 = head(c1, v1, (14, 14, 14), (112, 112, 112))
        y2 = head(c2, v2, (14, 14, 14), (112, 112, 112))
        y0 = head(torch.zeros_like(c1), torch.zeros_like(v1), (14, 14, 14), (112, 112, 112))
        y12 = head(c1 + c2, v1 + v2, (14, 14, 14), (112, 112, 112))
    assert y1.shape == (B, K, 112, 112, 112)
    assert rel_err(y12 - y0, (y1 - y0) + (y2 - y0)) < 1e-5
    # crop parity vs the oracle (reference formulation), full channel widths
    cs = torch.randn(1, Cf, 8, 8, 8)
    vs = torch.randn(1, 8, Fd)
    p = {"out_fpn_bridgeconv3d.weight": Wb, "out_fpn_bridgeconv3d.bias": bb, "out_conv3d.weight": Wc,
         "out_conv3d.bias": bc}
    ref = O.seg_head_3d(p, cs, vs, (2, 2, 2), (16, 16, 16), 2)
    with torch.no_grad():
        y = head(cs.cuda(), vs.cuda(), (2, 2, 2), (16, 16, 16))
    assert rel_err(y, ref) < 1e-4


@pytest.mark.parametrize("dims,A,grid,qkb", [
    ([1792, 1792, 896, 448], 256, (36, 36), False),      # cfg 2 stack (eff-b4, layercompress 1,1,2,2, --noqkbias), 36x36 grid
    ([2048, 2048, 2048], 256, (22, 22), True),           # cfg 3 stack (resnet50, 2 layers), quarter-size grid
])
def test_2d_config_stacks_match_oracle(dims, A, grid, qkb):
    """The 2-D BASELINE configs' channel widths (1792/896/448/2048, d = 448/224/512) at a reduced token grid."""
    import segtran_b200.networks.segtran_shared as S
    cfg = encoder_config(S.SegtranConfig, dims=dims, num_modes=4, num_attractors=A, pos_dim=2, qk_have_bias=qkb)
    cfg.translayer_compress_ratios = [1] * len(dims)
    torch.manual_seed(5)
    enc = S.SegtranFusionEncoder(cfg, "Fusion")
    init = S.SegtranInitWeights(cfg)
    enc.apply(init.init_weights)
    enc.apply(init.tie_qk)
    enc.apply(init.add_identity_bias)
    enc.eval()
    p = {"voxel_fusion." + k: v.clone() for k, v in enc.state_dict().items()}
    N = grid[0] * grid[1]
    x = torch.randn(2, N, dims[0])
    pos = O.voxels_pos_for_grid(grid, (8, 8), 2)
    mask = (torch.rand(2, N, 1) > 0.1).float()
    with torch.no_grad():
        ref = O.fusion_encoder(p, "voxel_fusion.", x, pos, mask, dims, 4)
        y = enc.cuda()(x.cuda(), pos.cuda(), mask.cuda(), torch.Size(grid))
    e = rel_err(y, ref)
    print("dims %s: max-rel %.3e rms-rel %.3e" % (dims, e, rms_rel(y, ref)))
    # Default (single-pass TF32) mode at the widest, deepest stacks: TF32 operand rounding ALONE puts this figure at
    # 0.7e-3 .. 1.4e-3 depending on the seed (CPU emulation of TF32 operands with fp64 accumulation,
    # tools/tf32_error_study.py), i.e. the 1e-3 line of the north star runs through the middle of the rounding noise
    # here.  The bound below is that noise floor; the 1e-3 guarantee at these widths is the tf32x3 assertion that follows
    # (and the single-pass mode keeps < 1e-3 with 2x margin on the cfg-4 stack the headline metric is quoted on).
    assert e < 2e-3
    # the same stack in the 3-pass validation mode: fp32-level agreement (TF32 rounding is the only deviation above)
    from segtran_b200 import ops
    ops.set_precision("tf32x3")
    try:
        with torch.no_grad():
            y3 = enc(x.cuda(), pos.cuda(), mask.cuda(), torch.Size(grid))
    finally:
        ops.set_precision("tf32")
    e3 = rel_err(y3, ref)
    print("dims %s tf32x3: max-rel %.3e" % (dims, e3))
    assert e3 < 5e-5
    # gradients flow and are finite at these widths
    xg = x.cuda().requires_grad_()
    enc.train()
    yg = enc(xg, pos.cuda(), mask.cuda(), torch.Size(grid))
    yg.square().mean().backward()
    assert torch.isfinite(xg.grad).all()
    assert all(torch.isfinite(q.grad).all() for q in enc.parameters() if q.grad is not None)
