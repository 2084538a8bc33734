#!/bin/bash
# conv0 on tensor cores vs the CUDA-core kernel: parity tests, then per-category step time of bench.py for both
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_path.py -q -m gpu -p no:cacheprovider -x -s -k "conv0 or hubert or c1 or batch_of" 2>&1 | grep -E "gate|passed|failed|Error|error" | tail -14
for m in tc simt; do echo "== AGB_CONV0=$m"; AGB_CONV0=$m python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(round(d['value'], 1), d['roofline']['category_ms_per_step'])"; done
