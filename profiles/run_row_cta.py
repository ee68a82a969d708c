"""The CTA-per-token row kernels (csrc/sx_rows_cta.cuh) at the cfg-4 and cfg-3 sizes: two rounds of
ln_softaggr fwd/bwd + prologue fwd/bwd (for `ncu -k regex:_cta -s 8 -c 8 --set full`), then CUDA-event timings."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from segtran_b200 import ops  # noqa: E402

torch.manual_seed(0)
dev = "cuda"
SIZES = [("cfg4", 4, 4, 2744, 1024), ("cfg3", 16, 4, 1936, 2048)]
seed = ops.new_dropout_seed(torch.device(dev, 0))


def make(B, M, N, F):
    Y = torch.randn(B, M, N, F, device=dev, requires_grad=True)
    g = torch.ones(F, device=dev, requires_grad=True)
    b = torch.zeros(F, device=dev, requires_grad=True)
    ws = (torch.randn(1, F, device=dev) * 0.02).requires_grad_()
    bs = torch.zeros(1, device=dev, requires_grad=True)
    x = torch.randn(B, N, F, device=dev, requires_grad=True)
    pe = torch.randn(N, F, device=dev)
    mask = torch.ones(B, N, 1, device=dev)
    return Y, g, b, ws, bs, x, pe, mask


def once(t, timed=None):
    Y, g, b, ws, bs, x, pe, mask = t
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    ev[0].record()
    out = ops.ln_softaggr(Y, g, b, ws, bs, 0.2, seed)
    ev[1].record()
    go = torch.ones_like(out)
    torch.cuda.synchronize()
    ev[1].record()
    out.backward(go)
    ev[2].record()
    torch.cuda.synchronize()
    g2 = torch.ones_like(g, requires_grad=True)
    ev[2].record()
    h = ops.prologue(x, g2, b, pe, 1.0, mask, 0.2, seed)
    ev[3].record()
    gh = torch.ones_like(h)
    torch.cuda.synchronize()
    ev[3].record()
    h.backward(gh)
    ev[4].record()
    torch.cuda.synchronize()
    return None


for name, B, M, N, F in SIZES:
    t = make(B, M, N, F)
    for _ in range(2):
        once(t)
torch.cuda.synchronize()
print("done")
