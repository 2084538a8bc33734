#!/bin/bash
# ncu evidence for profiles/: (1) every launch of one timed step with its device time and DRAM bytes, (2) full
# captures of the top kernels.  Each ncu command runs only after the identical plain command exited 0
# (B200_PROFILING.md).  One step of C2 = 376 launches; the first 3 (warm-up) steps are skipped.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
CMD_FULL="python bench.py --clips 64 --steps 1 --warmup 3 --no-cpu-baseline"
$CMD_FULL > gpurun_out/plain_full.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 1128 -c 376 --csv \
    --log-file gpurun_out/launches.csv $CMD_FULL > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 1 -c 6 -o gpurun_out/prof_gemm -f $CMD_FULL > gpurun_out/ncu_gemm.log 2>&1
echo "gemm capture rc=$?"
ncu --set full --clock-control none --import-source on -k regex:mha_tc2_kernel -s 0 -c 2 -o gpurun_out/prof_mha -f $CMD_FULL > gpurun_out/ncu_mha.log 2>&1
echo "mha capture rc=$?"
ncu --set full --clock-control none --import-source on -k "regex:ln_f32_kernel|conv0_kernel|ln512_bf16_kernel|embed_preln_kernel|patchify_kernel|seq_pool_kernel" -s 0 -c 6 -o gpurun_out/prof_hbm -f $CMD_FULL > gpurun_out/ncu_hbm.log 2>&1
echo "hbm capture rc=$?"
tail -2 gpurun_out/plain_full.log
ls -la gpurun_out | head -40
