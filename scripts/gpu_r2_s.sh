#!/bin/bash
cd "$(dirname "$0")/.."
cat > /tmp/t.py <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
from musicstyletransfer_b200 import ops
for (N, K) in ((768, 256), (1024, 256), (256, 1024)):
    M = 2048*65
    a = torch.randn(M, K, device="cuda"); w = torch.randn(N, K, device="cuda")*0.05; b = torch.randn(N, device="cuda"); y = torch.empty(M, N, device="cuda")
    for _ in range(3): ops.gemm_tc(a, K, 0, w, K, 1, y, N, M, N, K, bias=b, x3=True)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(20): ops.gemm_tc(a, K, 0, w, K, 1, y, N, M, N, K, bias=b, x3=True)
    e.record(); torch.cuda.synchronize()
    ref = a[:512].double() @ w.double().t() + b.double()
    err = float((y[:512].double() - ref).abs().max() / ref.abs().max())
    print("N=%d K=%d: %.1f us  err %.2e" % (N, K, s.elapsed_time(e)/20*1e3, err))
PY
python /tmp/t.py
