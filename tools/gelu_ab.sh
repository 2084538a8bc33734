#!/bin/bash
# A/B of the erf-GELU evaluation in the GEMM epilogue: tanh-identity form (default) vs Abramowitz-Stegun (exact to 1.5e-7)
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== tanh-form (default build)"; python tools/gemm_bench.py 2>&1 | grep -E "hub_ffn1|vit_fc1|hub_qkv"
python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -x -k "gemm" 2>&1 | tail -2
python -m pytest tests/test_gpu_path.py -q -m gpu -p no:cacheprovider -x -s 2>&1 | grep -E "gate|passed|failed" | tail -12
echo "== A&S exact"; AGB_NVCC_EXTRA=-DAGB_GEMM_GELU_EXACT python -m affectgpt_b200.build --force > /dev/null 2>&1
python tools/gemm_bench.py 2>&1 | grep -E "hub_ffn1|vit_fc1|hub_qkv"
python -m pytest tests/test_gpu_path.py -q -m gpu -p no:cacheprovider -x -s 2>&1 | grep -E "c1.*audio|passed|failed" | tail -4
