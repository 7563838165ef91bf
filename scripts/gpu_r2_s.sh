#!/bin/bash
cd "$(dirname "$0")/.."
cat > /tmp/t.py <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
from musicstyletransfer_b200 import ops
for (N, K) in ((768, 256), (1024, 256), (256, 1024), (256, 256), (293, 128), (128, 132)):
    M = 2048*65
    a = torch.randn(M, K, device="cuda"); w = torch.randn(N, K, device="cuda")*0.05; b = torch.randn(N, device="cuda"); ldc=(N+3)//4*4; y = torch.empty(M, ldc, device="cuda")
    for _ in range(3): ops.gemm_tc(a, K, 0, w, K, 1, y, ldc, M, N, K, bias=b, x3=True)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(20): ops.gemm_tc(a, K, 0, w, K, 1, y, ldc, M, N, K, bias=b, x3=True)
    e.record(); torch.cuda.synchronize()
    rows = torch.cat([torch.arange(0, 600), torch.arange(M-600, M)]).cuda()
    ref = a[rows].double() @ w.double().t() + b.double()
    err = float((y[rows][:, :N].double() - ref).abs().max() / ref.abs().max())
    print("BK=%s N=%d K=%d: %.1f us  err %.2e" % (os.environ.get("MSX_X3_BK"), N, K, s.elapsed_time(e)/20*1e3, err))
PY
python /tmp/t.py
timeout 600 python -m pytest tests/test_parity_bench_gpu.py tests/test_gemm_gpu.py -q -m gpu -k "gemm" 2>&1 | tail -3
