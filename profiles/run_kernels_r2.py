#!/usr/bin/env python
"""Driver for the round-2 `ncu --set full` captures: the 3xTF32 GEMM on the step's forward shapes, the single-pass TF32 GEMM
on the same shapes (for comparison), the attention forward with compensated scores, the A2 row kernels and the roll
front end, each launched a few times at the bench sizes (B = 2048, T = 65)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from musicstyletransfer_b200 import featurise, ops, synth  # noqa: E402

dev = "cuda"
M, D = 2048 * 65, 256
x = torch.randn(M, D, device=dev)
for N, K in ((768, 256), (1024, 256), (256, 1024)):
    a = torch.randn(M, K, device=dev)
    w = torch.randn(N, K, device=dev) * 0.05
    b = torch.randn(N, device=dev)
    y = torch.empty(M, N, device=dev)
    for x3 in (True, False):
        for _ in range(2):
            ops.gemm_tc(a, K, 0, w, K, 1, y, N, M, N, K, bias=b, x3=x3)
    for _ in range(2):
        ops.gemm_tc_b3(a, K, w, K, y, N, M, N, K, bias=b)
torch.cuda.synchronize()
B, T, H, dh = 2048, 65, 8, 32
qkv = torch.randn(B * T, 3 * H * dh, device=dev)
mask = torch.ones(B * T, device=dev)
ctx = torch.empty(B * T, H * dh, device=dev)
for x3 in (True, False):
    for _ in range(2):
        ops.attention_tc_fwd(qkv, mask, ctx, B, T, H, dh, x3_scores=x3)
torch.cuda.synchronize()
# A2 rows from 512 synthetic tracks, roll features
blobs, classes = synth.midi_files(n_files=512, ev_per_file=2048, seed=3)
order = np.argsort(np.asarray(classes), kind="stable")
soas = [featurise.parse_smf(blobs[i])[1][0] for i in order]
cls = np.asarray(classes, np.int32)[order]
tokens, n_tokens = featurise.tokenize_tracks_device(soas, dev)
rows = featurise.build_rows(tokens, n_tokens, torch.from_numpy(cls).to(dev),
                            torch.from_numpy(np.searchsorted(cls, np.arange(3)).astype(np.int32)).to(dev), 64)
idx = torch.arange(2048, dtype=torch.int32, device=dev)
for _ in range(2):
    featurise.gather_batch(rows, idx, 65)
roll = (torch.rand(2048, 64, 128, device=dev) < 0.04).to(torch.uint8)
renc = torch.empty(2048 * 65, 132, device=dev)
rdec = torch.empty(2048 * 64, 132, device=dev)
for _ in range(2):
    ops.roll_features(roll, renc, rdec, 2048, 64)
torch.cuda.synchronize()
print("ok")
