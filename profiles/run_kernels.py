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
