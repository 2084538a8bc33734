#!/bin/bash
# A/B of the two step-level changes: audio chain on a side stream, last ViT block pruned to the CLS row.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee gpurun_out/ab.log
tail -3 gpurun_out/pytest_gpu.log | tee -a gpurun_out/ab.log
for cfg in "1 0" "0 0" "1 1" "0 1"; do
  set -- $cfg
  echo "== AGB_OVERLAP=$1 AGB_VIT_DENSE_LAST=$2" | tee -a gpurun_out/ab.log
  AGB_OVERLAP=$1 AGB_VIT_DENSE_LAST=$2 timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_ov$1_dense$2.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['achieved'], d['clocks'])" | tee -a gpurun_out/ab.log
done
timeout 900 python tools/hf_gpu_compare.py 2>&1 | grep -v Warning | tail -8 | tee -a gpurun_out/ab.log
