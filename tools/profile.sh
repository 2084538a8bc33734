#!/bin/bash
# ncu evidence for profiles/: (1) per-launch device times of one timed step, (2) full captures of the top kernels.
# Each ncu command runs only after the identical plain command exited 0 (B200_PROFILING.md).
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
KREGEX='regex:gemm_kernel|mha_kernel|ln_f32_kernel|ln512_bf16_kernel|conv0_kernel|patchify_kernel|embed_preln_kernel|seq_pool_kernel|linear_kernel|attn_pool_kernel|mean_t_kernel|pair_sum_kernel|wave_partial_kernel|wave_finalize_kernel'
CMD_SMALL="python bench.py --clips 16 --steps 1 --warmup 3 --no-cpu-baseline"
CMD_FULL="python bench.py --clips 64 --steps 1 --warmup 3 --no-cpu-baseline"
$CMD_SMALL > gpurun_out/plain_small.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" -s 1125 -c 375 --csv \
    --log-file gpurun_out/launches.csv $CMD_SMALL > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
$CMD_FULL > gpurun_out/plain_full.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 1 -c 4 -o gpurun_out/prof_gemm -f $CMD_FULL > gpurun_out/ncu_gemm.log 2>&1
echo "gemm capture rc=$?"
ncu --set full --clock-control none --import-source on -k regex:mha_kernel -s 0 -c 1 -o gpurun_out/prof_mha -f $CMD_FULL > gpurun_out/ncu_mha.log 2>&1
echo "mha capture rc=$?"
ncu --set full --clock-control none --import-source on -k "regex:ln_f32_kernel|conv0_kernel|ln512_bf16_kernel" -s 0 -c 3 -o gpurun_out/prof_hbm -f $CMD_FULL > gpurun_out/ncu_hbm.log 2>&1
echo "hbm capture rc=$?"
tail -2 gpurun_out/plain_full.log
ls -la gpurun_out
