#!/bin/bash
# Staged GPU bring-up: each stage in its own process with its own timeout so that a trapped kernel (sticky CUDA
# error) or a hang costs one stage, not the call.  Logs land in gpurun_out/.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; echo "=== $name"; timeout "$STAGE_TIMEOUT" "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "$name rc=$rc"; tail -n "${TAILN:-15}" gpurun_out/$name.log; echo "$name $rc" >> gpurun_out/selftest_summary.txt; }
: > gpurun_out/selftest_summary.txt
STAGE_TIMEOUT=300 run gemm_diag python tools/gemm_diag.py
STAGE_TIMEOUT=600 run kernels_gemm python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "gemm" -p no:cacheprovider -s
STAGE_TIMEOUT=600 run kernels_conv python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "conv3 or posconv" -p no:cacheprovider -s
STAGE_TIMEOUT=600 run kernels_other python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "not gemm and not conv3 and not posconv" -p no:cacheprovider -s
STAGE_TIMEOUT=900 run path python -m pytest tests/test_gpu_path.py -q -m gpu -p no:cacheprovider -s
cat gpurun_out/selftest_summary.txt
if grep -q "^gemm_diag 0" gpurun_out/selftest_summary.txt && grep -q "^path 0" gpurun_out/selftest_summary.txt; then
  STAGE_TIMEOUT=900 TAILN=3 run bench python bench.py --steps 3 --warmup 3
  STAGE_TIMEOUT=300 TAILN=8 run smoke python __graft_entry__.py --smoke
fi
