#!/usr/bin/env python
"""Driver for the second batch of round-2 `ncu --set full` captures: the p3 GEMM (operands as bf16 hi / lo planes) on the
step's forward shapes, the single-pass TF32 GEMM on one of them for comparison, the attention kernels of the top layer
(q0_only) next to the general ones, and the LSTM recurrence (32-row and 16-row clusters), at the bench sizes."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from musicstyletransfer_b200 import ops  # noqa: E402

dev = "cuda"
REPS = int(os.environ.get("MSX_PROFILE_REPS", "2"))   # 1 under ncu (every launch is replayed ~40 times)
M = 2048 * 65
for N, K, relu in ((768, 256, False), (1024, 256, True), (256, 1024, False)):
    A = torch.randn(M, K, device=dev)
    W = torch.randn(N, K, device=dev) * 0.05
    b = torch.randn(N, device=dev)
    C = torch.empty(M, N, device=dev)
    Ah, Al = torch.empty_like(A, dtype=torch.bfloat16), torch.empty_like(A, dtype=torch.bfloat16)
    Wh, Wl = torch.empty_like(W, dtype=torch.bfloat16), torch.empty_like(W, dtype=torch.bfloat16)
    ops.split_planes(A, Ah, Al)
    ops.split_planes(W, Wh, Wl)
    mask = torch.zeros(M, N // 32, dtype=torch.int32, device=dev) if relu else None
    Ch, Cl = (torch.empty(M, N, device=dev, dtype=torch.bfloat16) for _ in range(2))
    for _ in range(REPS):
        if relu:
            ops.gemm_tc_p3(Ah, Al, K, Wh, Wl, K, Ch, N, M, N, K, bias=b, relu=True, drop_p=0.2, mask_out=mask, ldmask=N // 32, C_lo=Cl)
        else:
            ops.gemm_tc_p3(Ah, Al, K, Wh, Wl, K, C, N, M, N, K, bias=b)
    if N == 768:
        for _ in range(REPS):
            ops.gemm_tc(A, K, 0, W, K, 1, C, N, M, N, K, bias=b)
torch.cuda.synchronize()
B, T, H, dh = 2048, 65, 8, 32
D = H * dh
qkv = torch.randn(B * T, 3 * D, device=dev)
mask = torch.ones(B * T, device=dev)
hi = torch.empty(B * T, D, device=dev, dtype=torch.bfloat16)
lo = torch.empty_like(hi)
dctx = torch.randn(B * T, D, device=dev)
d0 = torch.zeros(B, T, D, device=dev)
d0[:, 0] = torch.randn(B, D, device=dev)
d0 = d0.view(B * T, D)
dqkv = torch.empty(B * T, 3 * D, device=dev)
db = torch.zeros(3 * D, device=dev)
for q0 in (False, True):
    for _ in range(REPS):
        ops.attention_tc_fwd(qkv, mask, hi, B, T, H, dh, x3_scores=True, ctx_lo=lo, q0_only=q0)
    for _ in range(REPS):
        ops.attention_tc_bwd(qkv, mask, d0 if q0 else dctx, dqkv, B, T, H, dh, dbias=db, q0_only=q0)
torch.cuda.synchronize()
Hd = 128
for Bl in (2048, 32):
    gx = torch.randn(Bl * T, 4 * Hd, device=dev) * 0.8
    w = torch.randn(4 * Hd, Hd, device=dev) * 0.12
    bh = torch.randn(4 * Hd, device=dev) * 0.1
    tv = torch.randn(Bl, 2 * Hd, device=dev) * 0.5
    dhs = torch.randn(Bl * T, Hd, device=dev) * 0.3
    hs, hp, cs = (torch.zeros(Bl * T, Hd, device=dev) for _ in range(3))
    dtv = torch.zeros(Bl, 2 * Hd, device=dev)
    for _ in range(REPS):
        ops.lstm_tc_fwd(gx, w, bh, tv, tv[:, Hd:], 2 * Hd, hs, hp, cs, Bl, T, Hd)
        ops.lstm_tc_bwd(gx, w, cs, tv[:, Hd:], 2 * Hd, dhs, dtv, dtv[:, Hd:], Bl, T, Hd)
torch.cuda.synchronize()
print("done")
