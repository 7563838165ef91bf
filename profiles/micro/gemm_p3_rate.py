"""Per-launch time of the forward GEMMs of the train step at the bench rows (M = 133 120) in the four product modes:
single-pass TF32, 3xTF32 (in-kernel split), bf16x3 (in-kernel split) and p3 (operands as bf16 hi / lo planes)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from musicstyletransfer_b200 import ops

M = 2048 * 65
for N, K, relu, cplanes in ((768, 256, False, False), (256, 256, False, False), (1024, 256, True, True), (256, 1024, False, False)):
    A = torch.randn(M, K, device="cuda")
    W = torch.randn(N, K, device="cuda") * 0.05
    b = torch.randn(N, device="cuda")
    C = torch.empty(M, N, device="cuda")
    mask = torch.zeros(M, N // 32, dtype=torch.int32, device="cuda") if relu else None
    Ah, Al = torch.empty_like(A, dtype=torch.bfloat16), torch.empty_like(A, dtype=torch.bfloat16)
    Wh, Wl = torch.empty_like(W, dtype=torch.bfloat16), torch.empty_like(W, dtype=torch.bfloat16)
    ops.split_planes(A, Ah, Al); ops.split_planes(W, Wh, Wl)
    Ch = torch.empty(M, N, device="cuda", dtype=torch.bfloat16); Cl = torch.empty_like(Ch)
    fns = {
        "tf32": lambda: ops.gemm_tc(A, K, 0, W, K, 1, C, N, M, N, K, bias=b, relu=relu, drop_p=0.2 if relu else 0.0, mask_out=mask, ldmask=N // 32),
        "tf32x3": lambda: ops.gemm_tc(A, K, 0, W, K, 1, C, N, M, N, K, bias=b, relu=relu, drop_p=0.2 if relu else 0.0, mask_out=mask, ldmask=N // 32, x3=True),
        "p3": lambda: ops.gemm_tc_p3(Ah, Al, K, Wh, Wl, K, C, N, M, N, K, bias=b, relu=relu, drop_p=0.2 if relu else 0.0, mask_out=mask, ldmask=N // 32),
    }
    if cplanes:
        fns["p3->planes"] = lambda: ops.gemm_tc_p3(Ah, Al, K, Wh, Wl, K, Ch, N, M, N, K, bias=b, relu=relu, drop_p=0.2, mask_out=mask, ldmask=N // 32, C_lo=Cl)
    for name, fn in fns.items():
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 100
        print("N=%4d K=%4d %-11s %7.1f us  %6.1f TFLOP/s (algorithmic)" % (N, K, name, us, 2.0 * M * N * K / us / 1e6))
