"""Forward error of the full-size BASELINE stacks (B=1) against the fp32 CPU oracle under different precision policies
of the default TF32 mode (ops.set_precision_policy).  usage: python profiles/run_precision_policy.py [cfgs] [seeds]"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import segtran_oracle as O
from tests.helpers import encoder_config, rel_err, rms_rel
import segtran_b200.networks.segtran_shared as S
from segtran_b200 import ops
import bench

cfgs = [int(c) for c in (sys.argv[1] if len(sys.argv) > 1 else "2,3,4").split(",")]
seeds = int(sys.argv[2]) if len(sys.argv) > 2 else 2
POL = {"all_tf32": dict(small="tf32", proj="tf32", insq="tf32"), "small_x3 (default)": dict(small="tf32x3", proj="tf32", insq="tf32"),
       "small+proj_x3": dict(small="tf32x3", proj="tf32x3", insq="tf32"),
       "small+proj+insq_x3": dict(small="tf32x3", proj="tf32x3", insq="tf32x3")}
torch.set_num_threads(min(32, os.cpu_count() or 8))
for ci in cfgs:
    c = bench.CONFIGS[ci]
    dims, A, grid = c["dims"], c["attractors"], c["grid"]
    pd = len(grid)
    for seed in range(seeds):
        cfg = encoder_config(S.SegtranConfig, dims=dims, num_modes=4, num_attractors=A, pos_dim=pd, qk_have_bias=c["qk_bias"])
        cfg.translayer_compress_ratios = [1] * len(dims)
        torch.manual_seed(100 + seed)
        enc = S.SegtranFusionEncoder(cfg, "Fusion")
        init = S.SegtranInitWeights(cfg)
        enc.apply(init.init_weights); enc.apply(init.tie_qk); enc.apply(init.add_identity_bias)
        enc.eval()
        p = {"voxel_fusion." + k: v.clone() for k, v in enc.state_dict().items()}
        N = 1
        for s in grid: N *= s
        x = torch.randn(1, N, dims[0])
        pos = O.voxels_pos_for_grid(grid, (8,) * pd, 1)
        mask = torch.ones(1, N, 1)
        t0 = time.time()
        with torch.no_grad():
            ref = O.fusion_encoder(p, "voxel_fusion.", x, pos, mask, dims, 4)
        enc = enc.cuda()
        out = []
        for name, pol in POL.items():
            ops.set_precision_policy(**pol)
            with torch.no_grad():
                y = enc(x.cuda(), pos.cuda(), mask.cuda(), torch.Size(grid))
            out.append("%s %.2e/%.2e" % (name, rel_err(y, ref), rms_rel(y, ref)))
        ops.set_precision_policy(small="tf32x3", proj="tf32", insq="tf32")
        print("cfg%d N=%d seed %d (oracle %.0fs): max-rel/rms-rel  " % (ci, N, seed, time.time() - t0) + " | ".join(out), flush=True)
