"""One launch of each long-row attention kernel (full and q0_only, forward and backward) at the L = 128 / 256 bench shapes'
token count, for ncu: ncu --set full -k regex:attn_tcl --launch-skip 4 --launch-count 4 python profiles/micro/prof_attn_long.py 129"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from musicstyletransfer_b200 import ops
H, dh, T = 8, 32, int(sys.argv[1]) if len(sys.argv) > 1 else 129
D = H * dh
B = 2048 * 129 // T
qkv = torch.randn(B * T, 3 * D, device="cuda"); mask = torch.ones(B * T, device="cuda"); dctx = torch.randn(B * T, D, device="cuda")
ctx = torch.empty(B * T, D, device="cuda"); dqkv = torch.empty(B * T, 3 * D, device="cuda"); db = torch.zeros(3 * D, device="cuda")
stats = torch.zeros(B * H * T, 2, device="cuda")
for _ in range(2):
    ops.attention_tcl_fwd(qkv, mask, ctx, stats, B, T, H, dh)
    ops.attention_tcl_bwd(qkv, mask, dctx, stats, dqkv, B, T, H, dh, dbias=db)
    ops.attention_tcl_fwd(qkv, mask, ctx, stats, B, T, H, dh, q0_only=True)
    ops.attention_tcl_bwd(qkv, mask, dctx, stats, dqkv, B, T, H, dh, dbias=db, q0_only=True)
torch.cuda.synchronize()
