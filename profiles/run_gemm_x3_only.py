#!/usr/bin/env python
"""ncu target: the 3xTF32 and the bf16x3 GEMM on the QKV forward shape only (source-level stall analysis)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from musicstyletransfer_b200 import ops  # noqa: E402

M, N, K = 2048 * 65, 768, 256
a = torch.randn(M, K, device="cuda")
w = torch.randn(N, K, device="cuda") * 0.05
b = torch.randn(N, device="cuda")
y = torch.empty(M, N, device="cuda")
for _ in range(2):
    ops.gemm_tc(a, K, 0, w, K, 1, y, N, M, N, K, bias=b, x3=True)
    ops.gemm_tc_b3(a, K, w, K, y, N, M, N, K, bias=b)
torch.cuda.synchronize()
print("ok")
