#!/usr/bin/env python
"""Small driver for `ncu --set full` captures: launches the K1 rasteriser on BASELINE config 2 and the three
tensor-GEMM modes on the train-vae shapes (B=2048, T=65) a few times each."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from musicstyletransfer_b200 import featurise, ops, synth  # noqa: E402

dev = "cuda"
d = [torch.from_numpy(a).to(dev) for a in synth.note_events()]
for _ in range(4):
    featurise.rasterize(*d)
M, D = 2048 * 65, 256
x = torch.randn(M, D, device=dev)
w = torch.randn(3 * D, D, device=dev)
b = torch.randn(3 * D, device=dev)
y = torch.empty(M, 3 * D, device=dev)
gw = torch.zeros(3 * D, D, device=dev)
dx = torch.empty(M, D, device=dev)
for _ in range(3):
    ops.gemm_tc(x, D, 0, w, D, 1, y, 3 * D, M, 3 * D, D, bias=b)                 # forward  [M,256] x [768,256]^T
    ops.gemm_tc(y, 3 * D, 0, w, D, 0, dx, D, M, D, 3 * D)                        # dgrad    [M,768] x [768,256]
    ops.gemm_tc(y, 3 * D, 1, x, D, 0, gw, D, 3 * D, D, M, splitk=24)             # wgrad    [M,768]^T x [M,256]
torch.cuda.synchronize()
print("ok")

# ---- attention and LSTM kernels at the train-vae shapes (B=2048, T=65, H=8, d_h=32; LSTM H=128)
B, T, H, dh = 2048, 65, 8, 32
qkv = torch.randn(B * T, 3 * H * dh, device=dev)
mask = torch.ones(B * T, device=dev)
ctx = torch.empty(B * T, H * dh, device=dev)
dctx = torch.randn(B * T, H * dh, device=dev)
dqkv = torch.empty_like(qkv)
db = torch.zeros(3 * H * dh, device=dev)
for _ in range(2):
    ops.attention_tc_fwd(qkv, mask, ctx, B, T, H, dh)
    ops.attention_tc_bwd(qkv, mask, dctx, dqkv, B, T, H, dh, dbias=db)
Hd = 128
gates = torch.randn(B * T, 4 * Hd, device=dev) * 0.1
w = torch.randn(4 * Hd, Hd, device=dev) * 0.05
bh = torch.zeros(4 * Hd, device=dev)
tv = torch.randn(B, 2 * Hd, device=dev) * 0.1
hs, hp, cs = (torch.empty(B * T, Hd, device=dev) for _ in range(3))
dhs = torch.randn(B * T, Hd, device=dev) * 0.1
dtv = torch.empty(B, 2 * Hd, device=dev)
dbi, dbh = torch.zeros(4 * Hd, device=dev), torch.zeros(4 * Hd, device=dev)
for _ in range(2):
    g2 = gates.clone()
    ops.lstm_fwd(g2, w, bh, tv, tv[:, Hd:], 2 * Hd, hs, hp, cs, B, T, Hd)
    ops.lstm_bwd(g2, w, cs, tv[:, Hd:], 2 * Hd, dhs, dtv, dtv[:, Hd:], B, T, Hd, db_i2h=dbi, db_h2h=dbh)
torch.cuda.synchronize()
print("ok2")
for _ in range(2):
    g2 = gates.clone()
    ops.lstm_tc_fwd(g2, w, bh, tv, tv[:, Hd:], 2 * Hd, hs, hp, cs, B, T, Hd)
    ops.lstm_tc_bwd(g2, w, cs, tv[:, Hd:], 2 * Hd, dhs, dtv, dtv[:, Hd:], B, T, Hd, db_i2h=dbi, db_h2h=dbh)
torch.cuda.synchronize()
print("ok3")
