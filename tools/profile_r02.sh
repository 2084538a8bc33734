#!/bin/bash
# ncu evidence of round 2 for profiles/ (run on the GPU box through gpurun; read back with tools/refresh_profiles.sh):
#  (1) every launch of one C2 step with device time and DRAM bytes (serialised, cold-cache: compare SHARES),
#  (2) --set full captures of the dominant GEMM launches, the attention kernel and the HBM-bound kernels,
#  (3) the tcgen05 GEMM next to cuBLAS on the path's shapes.
# Each ncu command runs only after the identical plain command exited 0 (B200_PROFILING.md).
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
CMD="python bench.py --clips 64 --steps 1 --warmup 3 --no-cpu-baseline --no-library-baseline --only-resident --no-graph"
KREGEX='regex:gemm_kernel|mha_tc2_kernel|mha_kernel|ln_f32_kernel|ln512_bf16_kernel|conv0_kernel|patchify_kernel|embed_preln_kernel|seq_pool_kernel|linear_kernel|attn_pool_kernel|mean_t_kernel|pair_sum_kernel|wave_partial_kernel|wave_finalize_kernel|gelu_bf16_kernel|concat2_kernel|conv0_tc_kernel|conv0_pack_kernel|cls_attention_kernel|cvt_kernel'
timeout 300 $CMD > gpurun_out/plain_full.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_full.log; exit 1; }
tail -1 gpurun_out/plain_full.log | cut -c1-200
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "$KREGEX" --csv \
    --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$? ($(grep -c gemm_kernel gpurun_out/launches.csv) gemm rows)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 1 -c 5 -o gpurun_out/prof_gemm -f $CMD > gpurun_out/ncu_gemm.log 2>&1
echo "gemm capture rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mha_tc2_kernel -s 0 -c 1 -o gpurun_out/prof_mha -f $CMD > gpurun_out/ncu_mha.log 2>&1
echo "mha capture rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:conv0_tc_kernel|ln_f32_kernel|ln512_bf16_kernel|patchify_kernel" -s 0 -c 4 -o gpurun_out/prof_hbm -f $CMD > gpurun_out/ncu_hbm.log 2>&1
echo "hbm capture rc=$?"
timeout 300 python tools/gemm_bench.py > gpurun_out/gemm_bench.log 2>&1; tail -10 gpurun_out/gemm_bench.log
ls -la gpurun_out | head -30
