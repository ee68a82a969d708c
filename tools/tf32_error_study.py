"""CPU study of where the TF32 operand-rounding error of the stack comes from (test infrastructure, not product).

Runs the oracle twice on the 2-D BASELINE stack widths: exact fp32, and with every matmul's operands rounded to TF32
(round-to-nearest-away, fp64 accumulate), optionally keeping selected call sites exact, over several seeds.
    python tools/tf32_error_study.py
"""
import sys, os
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import segtran_oracle as O
from tests.helpers import encoder_config, rel_err
import segtran_b200.networks.segtran_shared as S


def tf32(x):
    i = x.contiguous().view(torch.int32)
    i = (i + 0x1000) & ~0x1FFF
    return i.view(torch.float32)


EXACT = set()
COUNT = {}
LAYER = [0]


def _site(name):
    COUNT[name] = COUNT.get(name, 0) + 1
    return name in EXACT or ("L%d:*" % LAYER[0]) in EXACT or ("L%d:%s" % (LAYER[0], name)) in EXACT


class Patched:
    def __init__(self):
        self.lin, self.mm, self.es = F.linear, torch.matmul, torch.einsum

    def linear(self, x, w, b=None):
        site = "linear%dx%d" % (w.shape[0], w.shape[1])
        if _site(site):
            return self.lin(x, w, b)
        y = self.lin(tf32(x).double(), tf32(w).double()).float()
        return y if b is None else y + b

    def matmul(self, a, b):
        if _site("matmul"):
            return self.mm(a, b)
        return self.mm(tf32(a).double(), tf32(b).double()).float()

    def einsum(self, eq, a, b):
        if _site("einsum"):
            return self.es(eq, a, b)
        return self.es(eq, tf32(a).double(), tf32(b).double()).float()


def run(dims, A, grid, qkb, seed, emulate):
    cfg = encoder_config(S.SegtranConfig, dims=dims, num_modes=4, num_attractors=A, pos_dim=2, qk_have_bias=qkb)
    cfg.translayer_compress_ratios = [1] * len(dims)
    torch.manual_seed(seed)
    enc = S.SegtranFusionEncoder(cfg, "Fusion")
    init = S.SegtranInitWeights(cfg)
    enc.apply(init.init_weights); enc.apply(init.tie_qk); enc.apply(init.add_identity_bias)
    p = {"voxel_fusion." + k: v.clone() for k, v in enc.state_dict().items()}
    N = grid[0] * grid[1]
    x = torch.randn(2, N, dims[0])
    pos = O.voxels_pos_for_grid(grid, (8, 8), 2)
    mask = (torch.rand(2, N, 1) > 0.1).float()
    P = Patched()
    sq = O.squeezed_layer

    def sq_layer(p_, pre, *a, **k):
        LAYER[0] = int(pre.rstrip(".").split(".")[-1])
        return sq(p_, pre, *a, **k)

    O.squeezed_layer = sq_layer
    with torch.no_grad():
        ref = O.fusion_encoder(p, "voxel_fusion.", x, pos, mask, dims, 4)
        out = {}
        for name, exact in emulate.items():
            EXACT.clear(); EXACT.update(exact); COUNT.clear()
            O.F.linear, O.torch.matmul, O.torch.einsum = P.linear, P.matmul, P.einsum
            try:
                y = O.fusion_encoder(p, "voxel_fusion.", x, pos, mask, dims, 4)
            finally:
                O.F.linear, O.torch.matmul, O.torch.einsum = P.lin, P.mm, P.es
            out[name] = rel_err(y, ref)
    return out, dict(COUNT)


if __name__ == "__main__":
    torch.set_num_threads(32)
    cfgs = [([1792, 1792, 896, 448], 256, (36, 36), False), ([2048, 2048, 2048], 256, (22, 22), True)]
    for dims, A, grid, qkb in cfgs:
        for seed in range(5, 5 + int(sys.argv[1]) if len(sys.argv) > 1 else 8):
            L = len(dims) - 1
            variants = {"all_tf32": set(), "einsum_exact": {"einsum"}, "last_layer_exact": {"L%d:*" % (L - 1)},
                        "first_layer_exact": {"L0:*"}, "last_einsum_exact": {"L%d:einsum" % (L - 1)},
                        "all_linear_exact": {k for k in ("linear%dx%d" % (a, b) for a in (448, 896, 1792, 2048, 3584, 7168, 8192)
                                                                     for b in (448, 896, 1792, 2048))},
                        "einsum+matmul_exact": {"einsum", "matmul"}}
            o, c = run(dims, A, grid, qkb, seed, variants)
            print(dims, "seed", seed, {k: "%.2e" % v for k, v in o.items()}, flush=True)
        print(c)
