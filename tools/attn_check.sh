#!/bin/bash
# attention bring-up: parity tests, then the micro-benchmark with and without the SIMT stray-query path
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "attention" -p no:cacheprovider -x 2>&1 | tail -15 | tee gpurun_out/attn_tests.log
for m in 1 0; do echo "== AGB_ATTN_SIMT_ROW=$m"; AGB_ATTN_SIMT_ROW=$m timeout 300 python tools/attn_bench.py 2>&1 | tail -4 | tee gpurun_out/attn_bench_simt$m.log; done
