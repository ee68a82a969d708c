"""Times one out-FPN pyramid stage at the cfg-4 sizes (segtran3d.py:347-359, layer 2 -> 3): curr [4,480,56^3] -> conv1x1 -> 832
channels, + trilinear(higher [4,832,28^3]), GroupNorm(8) — fused (ops.fpn_stage) vs the stock PyTorch modules (cuDNN, TF32
convolutions allowed as PyTorch's default leaves them), forward + backward, CUDA events, median of 5."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from segtran_b200 import ops  # noqa: E402

torch.manual_seed(0)
dev = "cuda"
B, Cin, Cout, S = 4, 480, 832, 56
conv = torch.nn.Conv3d(Cin, Cout, 1).to(dev)
gn = torch.nn.GroupNorm(8, Cout).to(dev)
cur = torch.randn(B, Cin, S, S, S, device=dev, requires_grad=True)
hi = torch.randn(B, Cout, S // 2, S // 2, S // 2, device=dev, requires_grad=True)
G = torch.randn(B, Cout, S, S, S, device=dev)


def stock():
    up = conv(cur)
    y = gn(up + F.interpolate(hi, size=up.shape[2:], mode="trilinear", align_corners=False))
    (y * G).sum().backward()
    return y


def fused():
    y = ops.fpn_stage(cur, hi, conv, gn, "AN")
    (y * G).sum().backward()
    return y


def timeit(fn):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(5):
        for t in (cur, hi, conv.weight, conv.bias, gn.weight, gn.bias):
            t.grad = None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


with torch.no_grad():
    pass
ys = stock().detach()
yf = fused().detach()
print("fwd max-rel fused vs stock: %.2e" % float((yf - ys).abs().max() / ys.abs().max()))
ms_s, ms_f = timeit(stock), timeit(fused)
flops = 3 * 2.0 * B * Cin * Cout * S ** 3
print("stock  fwd+bwd %.2f ms   fused fwd+bwd %.2f ms   (conv GEMMs: %.0f GFLOP, %.0f TFLOP/s if they were all of it)" %
      (ms_s, ms_f, flops / 1e9, flops / ms_f / 1e9))
