"""Launches the HBM-bound kernels of the cfg-4 step once each (for `ncu --set full`)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from segtran_b200 import ops  # noqa: E402

torch.manual_seed(0)
dev = "cuda"
B, M, N, F = 4, 4, 2744, 1024
Y = torch.randn(B, M, N, F, device=dev, requires_grad=True)
g = torch.ones(F, device=dev, requires_grad=True)
b = torch.zeros(F, device=dev, requires_grad=True)
ws = (torch.randn(1, F, device=dev) * 0.02).requires_grad_()
bs = torch.zeros(1, device=dev, requires_grad=True)
seed = ops.new_dropout_seed(torch.device(dev, 0))
for _ in range(2):
    out = ops.ln_softaggr(Y, g, b, ws, bs, 0.2, seed)
    out.backward(torch.randn_like(out))
    S = torch.randn(B, M, N, 1024, device=dev, requires_grad=True)
    P = ops.softmax(S, None, 500.0, 0.2, seed)
    P.backward(torch.randn_like(P))
    S1 = torch.randn(B, 1, 1024, N, device=dev, requires_grad=True)
    P1 = ops.softmax(S1, None, 500.0, 0.2, seed)
    P1.backward(torch.randn_like(P1))
    curr = torch.randn(B, 832, 56, 56, 56, device=dev, requires_grad=True)
    vf = torch.randn(B, N, F, device=dev, requires_grad=True)
    Wb = (torch.randn(F, 832, 1, 1, 1, device=dev) * 0.02).requires_grad_()
    bb = torch.zeros(F, device=dev, requires_grad=True)
    Wc = (torch.randn(4, F, 1, 1, 1, device=dev) * 0.02).requires_grad_()
    bc = torch.zeros(4, device=dev, requires_grad=True)
    lo = ops.seg_head(curr, vf, (14, 14, 14), Wb, bb, Wc, bc, (112, 112, 112), d_pool_k=2)
    lo.backward(torch.randn_like(lo))
torch.cuda.synchronize()
print("done")
