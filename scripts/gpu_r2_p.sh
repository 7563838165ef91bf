#!/bin/bash
cd "$(dirname "$0")/.."
cat > /tmp/t.py <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
from musicstyletransfer_b200 import ops
for B in (2048, 32):
    T, Hd = 65, 128
    gates = torch.randn(B * T, 4 * Hd, device="cuda") * 0.1
    w = torch.randn(4 * Hd, Hd, device="cuda") * 0.05
    bh = torch.zeros(4 * Hd, device="cuda")
    tv = torch.randn(B, 2 * Hd, device="cuda") * 0.1
    hs, hp, cs = (torch.empty(B * T, Hd, device="cuda") for _ in range(3))
    g2 = gates.clone()
    for _ in range(3): ops.lstm_tc_fwd(g2, w, bh, tv, tv[:, Hd:], 2 * Hd, hs, hp, cs, B, T, Hd)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10): ops.lstm_tc_fwd(g2, w, bh, tv, tv[:, Hd:], 2 * Hd, hs, hp, cs, B, T, Hd)
    e.record(); torch.cuda.synchronize()
    print("DEBUG=%s B=%d: %.1f us (%.2f us/step)" % (os.environ.get("MSX_LSTM_DEBUG"), B, s.elapsed_time(e)/10*1e3, s.elapsed_time(e)/10*1e3/T))
PY
MSX_LSTM_DEBUG=0 python /tmp/t.py; timeout 600 python -m pytest tests/test_lstm_gpu.py tests/test_parity_bench_gpu.py tests/test_modules_gpu.py tests/test_engine_gpu.py -q -m gpu -k "lstm or style or beam or stacked" 2>&1 | tail -4
