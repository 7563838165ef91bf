#!/usr/bin/env python
"""Per-shape timing of the tensor GEMM kernels on the GEMM shapes of one train-vae step (B=2048, T=65):
1-CTA 128x128 tiles vs CTA-pair (cta_group::2) 256x256 tiles.  CUDA events, operands rotated over buffers
larger than L2.  Prints one line per (shape, kernel): us, TFLOP/s, algorithmic GB/s (fp32 A + B + C once)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from musicstyletransfer_b200 import ops  # noqa: E402

dev = "cuda"
M = 2048 * 65
D, F, H, V = 256, 1024, 128, 296
SHAPES = [
    # name, mode, (M, N, K) as the math sees them
    ("qkv fwd", "fwd", M, 3 * D, D), ("proj fwd", "fwd", M, D, D), ("ff1 fwd", "fwd", M, F, D), ("ff2 fwd", "fwd", M, D, F),
    ("i2h fwd", "fwd", M, 4 * H, H), ("out fwd", "fwd", M, V, H),
    ("ff2 dgrad", "dgrad", M, F, D), ("ff1 dgrad", "dgrad", M, D, F), ("qkv dgrad", "dgrad", M, D, 3 * D),
    ("out dgrad", "dgrad", M, H, V), ("i2h dgrad", "dgrad", M, H, 4 * H),
    ("qkv wgrad", "wgrad", 3 * D, D, M), ("ff1 wgrad", "wgrad", F, D, M), ("ff2 wgrad", "wgrad", D, F, M),
    ("proj wgrad", "wgrad", D, D, M), ("i2h wgrad", "wgrad", 4 * H, H, M), ("out wgrad", "wgrad", V, H, M),
]
NBUF = 3


def run(name, mode, m, n, k, pair):
    ops.gemm_tc_set_pair(pair)
    if mode == "fwd":
        ta, tb = 0, 1
        A = [torch.randn(m, k, device=dev) for _ in range(NBUF)]
        B = torch.randn(n, k, device=dev)
    elif mode == "dgrad":
        ta, tb = 0, 0
        A = [torch.randn(m, k, device=dev) for _ in range(NBUF)]
        B = torch.randn(k, n, device=dev)
    else:
        ta, tb = 1, 0
        A = [torch.randn(k, m, device=dev) for _ in range(NBUF)]
        B = [torch.randn(k, n, device=dev) for _ in range(NBUF)]
    C = [torch.zeros(m, n, device=dev) for _ in range(NBUF if mode != "wgrad" else 1)]
    sk = max(ops.wgrad_splitk(m, n, k), 2) if mode == "wgrad" else 1

    def call(i):
        a = A[i % NBUF]
        b = B[i % NBUF] if isinstance(B, list) else B
        c = C[i % len(C)]
        ops.gemm_tc(a, a.shape[1], ta, b, b.shape[1], tb, c, n, m, n, k, splitk=sk)

    for i in range(3):
        call(i)
    torch.cuda.synchronize()
    reps = 12
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for i in range(reps):
        evs[i][0].record()
        call(i)
        evs[i][1].record()
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in evs)[reps // 2]
    flops = 2.0 * m * n * k
    byts = 4.0 * (m * k + n * k + m * n)
    print("%-11s %-5s pair=%d  %8.1f us  %7.1f TFLOP/s  %7.1f GB/s" % (name, mode, pair, ms * 1e3, flops / ms / 1e9,
                                                                     byts / ms / 1e6), flush=True)
    return ms


tot = {0: 0.0, 1: 0.0}
for name, mode, m, n, k in SHAPES:
    for pair in (0, 1):
        tot[pair] += run(name, mode, m, n, k, pair)
print("sum over shapes: 1-CTA %.3f ms, pair %.3f ms" % (tot[0], tot[1]))
