#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "=== new tests"; timeout 900 python -m pytest tests/test_rows_gpu.py tests/test_smf_gpu.py tests/test_sanitize_gpu.py tests/test_engine_gpu.py -q -m gpu 2>&1 | tail -25
echo "=== sanitizer"; timeout 2400 bash scripts/sanitize.sh r2 2>&1 | tail -40
