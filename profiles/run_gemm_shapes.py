"""Launches the three GEMM shapes that dominate the cfg-4 step (for `ncu --set full -k regex:sx_gemm_kernel`):
  1. P.V of the squeeze-out attention   [16 x (2744 x 1024 x 1024)]  plain epilogue
  2. P.V' with MMSharedMid's epilogue    [16 x (2744 x 1024 x 1024)]  bias, pre-activation store, erf-GELU, dropout mask
  3. Q.K^T scores of the squeeze-out     [16 x (2744 x 1024 x 256)]   short K, fp32 scores + running max
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from segtran_b200 import ops  # noqa: E402

torch.manual_seed(0)
dev = "cuda"
P = torch.rand(4, 4, 2744, 1024, device=dev)
V = torch.randn(4, 1024, 4096, device=dev)
U = torch.randn(4, 4, 2744, 1024, device=dev)
W = torch.randn(1024, 1024, device=dev) * 0.02
b = torch.zeros(1024, device=dev)
q = torch.randn(4, 2744, 1024, device=dev)
k = torch.randn(4, 1024, 1024, device=dev)
amax = torch.full((1,), -3e38, device=dev)
for _ in range(2):
    vv = V.view(4, 1024, 4, 1024).permute(0, 2, 3, 1)
    out1 = ops.gemm_nt(P, vv)
    y = torch.empty(4, 4, 2744, 1024, device=dev)
    h = torch.empty_like(y)
    ops.gemm_nt(P, vv, out=y, bias=b, gelu=True, preact=h, drop_p=0.2, seed=1234)
    qv = q.view(4, 2744, 4, 256).permute(0, 2, 1, 3)
    kv = k.view(4, 1024, 4, 256).permute(0, 2, 1, 3)
    S = torch.empty(4, 4, 2744, 1024, device=dev)
    ops.gemm_nt(qv, kv, out=S, alpha=1 / 16.0, amax=amax, round_out=False)
torch.cuda.synchronize()
print("done")
