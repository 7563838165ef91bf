"""What the epilogue features of the two epilogue-paced GEMMs of the step cost (B = 2048, T = 65: M = 133 120):
FF1 forward (p3 planes in, N = 1024, K = 256: bias + ReLU + dropout + ReLU/dropout bit mask + hi/lo plane split) and
FF2 dgrad (tf32, N = 1024, K = 256: bit-mask aux + bias-gradient column sums), one feature switched off at a time."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from musicstyletransfer_b200 import ops

M, N, K = 2048 * 65, 1024, 256


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1000 / n


A = torch.randn(M, K, device="cuda")
W = torch.randn(N, K, device="cuda") * 0.05
b = torch.randn(N, device="cuda")
C = torch.empty(M, N, device="cuda")
mask = torch.zeros(M, N // 32, dtype=torch.int32, device="cuda")
Ah, Al = torch.empty_like(A, dtype=torch.bfloat16), torch.empty_like(A, dtype=torch.bfloat16)
Wh, Wl = torch.empty_like(W, dtype=torch.bfloat16), torch.empty_like(W, dtype=torch.bfloat16)
ops.split_planes(A, Ah, Al); ops.split_planes(W, Wh, Wl)
Ch = torch.empty(M, N, device="cuda", dtype=torch.bfloat16); Cl = torch.empty_like(Ch)
print("FF1 forward, p3 planes -> planes")
for name, kw, planes in (
        ("full (bias relu drop mask planes)", dict(bias=b, relu=True, drop_p=0.2, mask_out=mask, ldmask=N // 32), True),
        ("no dropout", dict(bias=b, relu=True, drop_p=0.0, mask_out=mask, ldmask=N // 32), True),
        ("no mask", dict(bias=b, relu=True, drop_p=0.2), True),
        ("no dropout, no mask", dict(bias=b, relu=True), True),
        ("no bias relu dropout mask", dict(), True),
        ("full, fp32 C", dict(bias=b, relu=True, drop_p=0.2, mask_out=mask, ldmask=N // 32), False),
        ("plain, fp32 C", dict(), False)):
    if planes:
        us = timeit(lambda: ops.gemm_tc_p3(Ah, Al, K, Wh, Wl, K, Ch, N, M, N, K, C_lo=Cl, **kw))
    else:
        us = timeit(lambda: ops.gemm_tc_p3(Ah, Al, K, Wh, Wl, K, C, N, M, N, K, **kw))
    print("  %-36s %7.1f us" % (name, us))
# FF2 dgrad: dh [M, 1024] = df [M, 256] @ W2 [256, 1024], masked by the FF1 bit mask, column sums -> ff1 bias gradient
dF = torch.randn(M, 256, device="cuda")
W2 = torch.randn(256, 1024, device="cuda") * 0.05
dH = torch.empty(M, 1024, device="cuda")
cs = torch.zeros(1024, device="cuda")
mask.random_(-2 ** 31, 2 ** 31 - 1)
print("FF2 dgrad, tf32")
for name, kw in (("full (bit-mask aux, colsum)", dict(aux=mask, ldaux=N // 32, aux_scale=1.25, out_colsum=cs)),
                 ("no colsum", dict(aux=mask, ldaux=N // 32, aux_scale=1.25)),
                 ("no aux", dict(out_colsum=cs)),
                 ("plain", dict())):
    us = timeit(lambda: ops.gemm_tc(dF, 256, 0, W2, 1024, 0, dH, 1024, M, 1024, 256, **kw))
    print("  %-36s %7.1f us" % (name, us))
# QKV-like plain shapes for reference
for n_, k_ in ((768, 256), (256, 1024)):
    A2 = torch.randn(M, k_, device="cuda"); W_ = torch.randn(n_, k_, device="cuda") * 0.05; C2 = torch.empty(M, n_, device="cuda")
    us = timeit(lambda: ops.gemm_tc(A2, k_, 0, W_, k_, 1, C2, n_, M, n_, k_))
    print("tf32 plain N=%d K=%d %7.1f us" % (n_, k_, us))
