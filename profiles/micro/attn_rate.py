"""Per-launch time of the tensor-core attention kernels at the bench shape (B = 2048, T = 65, H = 8, d_h = 32)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from musicstyletransfer_b200 import ops

B, T, H, dh = 2048, 65, 8, 32
D = H * dh
qkv = torch.randn(B * T, 3 * D, device="cuda")
mask = torch.ones(B * T, device="cuda")
ctx = torch.empty(B * T, D, device="cuda")
hi = torch.empty(B * T, D, device="cuda", dtype=torch.bfloat16); lo = torch.empty_like(hi)
dctx = torch.randn(B * T, D, device="cuda")
dqkv = torch.empty(B * T, 3 * D, device="cuda")
dctx0 = torch.zeros(B, T, D, device="cuda"); dctx0[:, 0] = torch.randn(B, D, device="cuda"); dctx0 = dctx0.view(B * T, D)
fns = {
    "fwd fp32 out": lambda: ops.attention_tc_fwd(qkv, mask, ctx, B, T, H, dh),
    "fwd x3 fp32 out": lambda: ops.attention_tc_fwd(qkv, mask, ctx, B, T, H, dh, x3_scores=True),
    "fwd x3 planes": lambda: ops.attention_tc_fwd(qkv, mask, hi, B, T, H, dh, x3_scores=True, ctx_lo=lo),
    "fwd x3 planes q0": lambda: ops.attention_tc_fwd(qkv, mask, hi, B, T, H, dh, x3_scores=True, ctx_lo=lo, q0_only=True),
    "fwd q0 fp32": lambda: ops.attention_tc_fwd(qkv, mask, ctx, B, T, H, dh, q0_only=True),
    "bwd": lambda: ops.attention_tc_bwd(qkv, mask, dctx, dqkv, B, T, H, dh),
    "bwd q0": lambda: ops.attention_tc_bwd(qkv, mask, dctx0, dqkv, B, T, H, dh, q0_only=True),
}
for name, fn in fns.items():
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record(); torch.cuda.synchronize()
    print("%-20s %7.1f us" % (name, e0.elapsed_time(e1) * 100))
