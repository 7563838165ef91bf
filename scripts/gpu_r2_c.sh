#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "=== targeted tests"; timeout 900 python -m pytest tests/test_attention_gpu.py tests/test_engine_gpu.py -q -m gpu -s -k "compensated or sos_rows" 2>&1 | grep -E "attention forward|passed|failed|FAILED|Error|assert" | head -40
echo "=== precision diag"; timeout 900 python profiles/micro/diag_precision.py 2>&1 | tail -12
