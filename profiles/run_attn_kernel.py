"""Times the fused scores+softmax kernel (sx_attn_probs_fwd) alone at the squeeze-out shapes of the BASELINE configs.
usage: python profiles/run_attn_kernel.py [cfg4|cfg2|cfg5] [iters]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from segtran_b200 import ops

SHAPES = {"cfg4": (4, 4, 2744, 1024, 256), "cfg2": (6, 4, 5184, 256, 448), "cfg5": (2, 4, 5832, 2048, 256),
          "cfg3": (16, 4, 1936, 256, 512)}
name = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
if len(sys.argv) > 3:                       # 1 = one launch (row-block items), 2 = two launches over tiles
    from segtran_b200 import _lib
    _lib.call("sx_gemm_debug_set", b"attn_mode", int(sys.argv[3]))
if len(sys.argv) > 4:
    from segtran_b200 import _lib
    _lib.call("sx_gemm_debug_set", b"attn_dbg", int(sys.argv[4]))
B, M, U1, U2, d = SHAPES[name]
q = torch.randn(B, U1, M * d, device="cuda")
k = torch.randn(B, U2, M * d, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for need_s, p in ((False, 0.0), (True, 0.0), (False, 0.2), (True, 0.2)):
    for _ in range(3):
        ops.attn_probs_fused(q, k, M, drop_p=p, seed=1, need_scores=need_s)
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.attn_probs_fused(q, k, M, drop_p=p, seed=1, need_scores=need_s)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    ms = ts[len(ts) // 2]
    passes = 2 if U2 > 256 else 1
    fl = 2.0 * B * M * U1 * U2 * d * passes
    print("%s store_S=%d dropout=%.1f: %.1f us  executed %.0f TFLOP/s  (P bytes %.0f MB)" % (
        name, need_s, p, ms * 1e3, fl / ms / 1e9, B * M * U1 * U2 * 4 / 1e6))
