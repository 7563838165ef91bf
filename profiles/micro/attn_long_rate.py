"""Per-launch time of the long-row attention kernels (attention_tc_long.cu: 128 < T <= 384) at a constant token count
(B * T ~ 264 k rows = the L = 128 bench step), next to the pipelined short-row kernels at T = 65 / 128.
Run on the GPU box: python profiles/micro/attn_long_rate.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from musicstyletransfer_b200 import ops

H, dh = 8, 32
D = H * dh
TOK = 2048 * 129


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1000 / n


for T in [int(a) for a in sys.argv[1:]] or (65, 128, 129, 130, 132, 133, 193, 257, 260, 384, 513, 768):
    B = max(1, TOK // T)
    qkv = torch.randn(B * T, 3 * D, device="cuda")
    mask = torch.ones(B * T, device="cuda")
    dctx = torch.randn(B * T, D, device="cuda")
    ctx = torch.empty(B * T, D, device="cuda")
    dqkv = torch.empty(B * T, 3 * D, device="cuda")
    db = torch.zeros(3 * D, device="cuda")
    if T <= 128:
        f = timed(lambda: ops.attention_tc_fwd(qkv, mask, ctx, B, T, H, dh))
        b = timed(lambda: ops.attention_tc_bwd(qkv, mask, dctx, dqkv, B, T, H, dh, dbias=db))
        name = "attn_tc "
    else:
        stats = torch.zeros(B * H * T, 2, device="cuda")
        f = timed(lambda: ops.attention_tcl_fwd(qkv, mask, ctx, stats, B, T, H, dh))
        b = timed(lambda: ops.attention_tcl_bwd(qkv, mask, dctx, stats, dqkv, B, T, H, dh, dbias=db))
        name = "attn_tcl"
        f0 = timed(lambda: ops.attention_tcl_fwd(qkv, mask, ctx, stats, B, T, H, dh, q0_only=True))
        b0 = timed(lambda: ops.attention_tcl_bwd(qkv, mask, dctx, stats, dqkv, B, T, H, dh, dbias=db, q0_only=True))
        print("attn_tcl T=%3d B=%4d: q0_only fwd %7.1f us   bwd %7.1f us" % (T, B, f0, b0))
    if T > 384:                                   # what ran these rows before the two-sweep forward: the exact FFMA key-tiled kernels
        ft = timed(lambda: ops.attention_tiled_fwd(qkv, mask, ctx, B, T, H, dh), n=3)
        bt = timed(lambda: ops.attention_tiled_bwd(qkv, mask, dctx, dqkv, B, T, H, dh), n=3)
        print("attn_tiled (FFMA) T=%3d B=%4d: fwd %7.1f us   bwd %7.1f us" % (T, B, ft, bt))
    items = B * H
    print("%s T=%3d B=%4d: fwd %7.1f us (%5.2f us per (b,h) per SM)   bwd %7.1f us (%5.2f)   fwd %.1f / bwd %.1f ns per token"
          % (name, T, B, f, f * 148 / items, b, b * 148 / items, f * 1e3 / (B * T), b * 1e3 / (B * T)))
