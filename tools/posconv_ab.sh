#!/bin/bash
# positional-conv bring-up: parity test + timing for the two A-operand variants
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for m in slab tap; do
  echo "== AGB_POSCONV=$m"
  AGB_POSCONV=$m timeout 180 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "posconv" -p no:cacheprovider -x 2>&1 | grep -E "passed|failed|Error|rel|assert" | head -6
  AGB_POSCONV=$m timeout 180 python tools/gemm_bench.py 2>&1 | grep posconv
done 2>&1 | tee gpurun_out/posconv_ab.log
