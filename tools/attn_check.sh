#!/bin/bash
# attention bring-up: parity tests, then the micro-benchmark: odd query row in the softmax warps (default) vs as a third tile
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "attention" -p no:cacheprovider -x 2>&1 | tail -5 | tee gpurun_out/attn_tests.log
for m in 0 1; do echo "== AGB_ATTN_STRAY_TILE=$m"; AGB_ATTN_STRAY_TILE=$m timeout 300 python tools/attn_bench.py 2>&1 | tail -2 | tee gpurun_out/attn_bench_tile$m.log; done
AGB_ATTN_TRACE=1 timeout 120 python tools/attn_bench.py 2>&1 | grep attn-trace | head -14 | tee gpurun_out/attn_trace.log
