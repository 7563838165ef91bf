#!/bin/bash
# compute-sanitizer over the kernel parity tests (run on the GPU box through gpurun):
#   scripts/sanitize.sh [tag]
# memcheck on every kernel test file, racecheck + synccheck on the kernels with hand-rolled shared-memory / mbarrier
# protocols.  tcgen05 / TMA kernels run under memcheck and synccheck; racecheck does not model the async proxy, so it is
# applied to the SIMT kernels (rasteriser, LayerNorm, losses, embedding, FFMA GEMM / attention / LSTM, Adam, beam, sample).
# Each tool's summary line lands in gpurun_out/sanitize_<tag>.txt (copy it to profiles/ to have it judged).
tag=${1:-run}
out=gpurun_out/sanitize_$tag.txt
mkdir -p gpurun_out
: > $out
SAN=${SAN:-compute-sanitizer}
# small, fast selections: the sanitizer slows kernels down 10-100x
MEM_TESTS="tests/test_rasterize_gpu.py::test_ragged_empty_and_long_sequences tests/test_rasterize_gpu.py::test_malformed_data_bytes_are_masked \
tests/test_layernorm_gpu.py tests/test_loss_gpu.py tests/test_embed_gpu.py tests/test_lstm_gpu.py tests/test_rows_gpu.py \
tests/test_smf_gpu.py tests/test_sanitize_gpu.py"
RACE_TESTS="tests/test_rasterize_gpu.py::test_ragged_empty_and_long_sequences tests/test_layernorm_gpu.py tests/test_loss_gpu.py \
tests/test_embed_gpu.py tests/test_rows_gpu.py tests/test_sanitize_gpu.py"
run() {   # tool, tests...
  tool=$1; shift
  echo "=== $tool: $*" >> $out
  MSX_SANITIZE=1 timeout 1500 $SAN --tool $tool --error-exitcode 86 --print-limit 20 \
    python -m pytest -x -q -m gpu -p no:cacheprovider $* > gpurun_out/sanitize_${tag}_$tool.log 2>&1
  rc=$?
  echo "exit code $rc" >> $out
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|error" gpurun_out/sanitize_${tag}_$tool.log | tail -8 >> $out
}
existing() { for t in "$@"; do f=${t%%::*}; [ -f "$f" ] && echo -n "$t "; done; }
run memcheck $(existing $MEM_TESTS)
run racecheck $(existing $RACE_TESTS)
run synccheck $(existing $MEM_TESTS)
cat $out
