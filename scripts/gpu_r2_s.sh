#!/bin/bash
cd "$(dirname "$0")/.."
cat > /tmp/t.py <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
from musicstyletransfer_b200 import ops
x3 = os.environ.get("X3") == "1"
for (N, K) in ((768, 256), (1024, 256), (256, 1024)):
    M = 2048*65
    a = torch.randn(M, K, device="cuda"); w = torch.randn(N, K, device="cuda")*0.05; b = torch.randn(N, device="cuda"); y = torch.empty(M, N, device="cuda")
    for _ in range(3): ops.gemm_tc(a, K, 0, w, K, 1, y, N, M, N, K, bias=b, x3=x3)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(20): ops.gemm_tc(a, K, 0, w, K, 1, y, N, M, N, K, bias=b, x3=x3)
    e.record(); torch.cuda.synchronize()
    print("x3=%s EPI=%s DEBUG=%s N=%d K=%d: %.1f us" % (x3, os.environ.get("MSX_GEMM_EPI_WARPS"), os.environ.get("MSX_X3_DEBUG"), N, K, s.elapsed_time(e)/20*1e3))
PY
python /tmp/t.py; MSX_GEMM_EPI_WARPS=8 python /tmp/t.py; MSX_GEMM_EPI_WARPS=16 python /tmp/t.py; X3=1 MSX_X3_DEBUG=3 python /tmp/t.py; X3=1 MSX_X3_DEBUG=1 python /tmp/t.py
