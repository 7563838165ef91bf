#!/bin/bash
cd "$(dirname "$0")/.."
echo "=== lstm tests"; timeout 600 python -m pytest tests/test_lstm_gpu.py tests/test_parity_bench_gpu.py tests/test_modules_gpu.py -q -m gpu -k "lstm or style or beam" 2>&1 | tail -4
MSX_LSTM_DEBUG=0 python /tmp/t.py 2>/dev/null || true
