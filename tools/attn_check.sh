#!/bin/bash
# attention bring-up: parity tests, then the micro-benchmark for each kernel generation (separate processes)
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "attention" -p no:cacheprovider -x 2>&1 | tail -15 | tee gpurun_out/attn_tests.log
for m in tc2 hmma; do echo "== $m"; AGB_ATTENTION=$m timeout 300 python tools/attn_bench.py 2>&1 | tail -4 | tee gpurun_out/attn_bench_$m.log; done
