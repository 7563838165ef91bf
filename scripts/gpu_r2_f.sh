#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "=== roll + engine tests"; timeout 900 python -m pytest tests/test_roll_gpu.py tests/test_engine_gpu.py tests/test_modules_gpu.py -q -m gpu -s 2>&1 | grep -E "roll step|roll bce|passed|failed|FAILED|Error|assert " | head -40
